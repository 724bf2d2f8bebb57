function [u_opt, x_opt, exitflag, fval, slack_opt] = ltvmpc_sqp_b200(model, n_sqp, x0, x_ref, dt, x_lin, u_lin)
%LTVMPC_SQP_B200 n_sqp repeated relinearise + condense + QP passes on a frozen x0 / x_ref: what main.m:118-127
%does across consecutive time steps (x_lin / u_lin = the previous x_opt / u_opt), iterated within one step
%(BASELINE configs[3]).  model: 0 kinematic, 1 dynamic.  Batched like ltvmpc_kinetmatic_curvilinear_b200.
    h = fsae_mpc_b200_handle();
    [u_opt, x_opt, exitflag, fval, slack_opt] = fsae_mpc_b200_mex('sqp', h, model, n_sqp, x0, x_ref, dt, x_lin, u_lin);
    exitflag = double(exitflag);
end
