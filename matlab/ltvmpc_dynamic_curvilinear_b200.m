function [u_opt, x_opt, QP, exitflag, fval, slack_opt] = ltvmpc_dynamic_curvilinear_b200(x0, x_ref, kappa, dt, x_lin, u_lin, QP)
%LTVMPC_DYNAMIC_CURVILINEAR_B200 Drop-in for mpc/ltv/dynamic/ltvmpc_dynamic_curvilinear.m
%(same contract as ltvmpc_kinetmatic_curvilinear_b200, 7 states, 4 slack variables).

    h = fsae_mpc_b200_handle();
    if isstruct(kappa)
        fsae_mpc_b200_mex('set_track', h, 0, kappa.x_spline, kappa.y_spline, kappa.dl);
    end
    [u_opt, x_opt, exitflag, fval, slack_opt] = fsae_mpc_b200_mex('ltvmpc', h, 1, x0, x_ref, dt, x_lin, u_lin);
    exitflag = double(exitflag);
    if any(exitflag)
        display(exitflag)      % ltvmpc_dynamic_curvilinear.m:53-55
    end
end
