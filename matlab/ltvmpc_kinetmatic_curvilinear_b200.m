function [u_opt, x_opt, QP, exitflag, fval, slack_opt] = ltvmpc_kinetmatic_curvilinear_b200(x0, x_ref, kappa, dt, x_lin, u_lin, QP)
%LTVMPC_KINETMATIC_CURVILINEAR_B200 Drop-in for mpc/ltv/kinematic/ltvmpc_kinetmatic_curvilinear.m
%that runs the whole step (linearise, condense, QP solve) on a B200 through the
%fsae_mpc_b200 MEX gateway.  Same arguments, same outputs; works for one problem
%(x0 [5x1], x_ref [5xN]) or a batch (x0 [5xB], x_ref/x_lin [5xNxB], u_lin [2xNxB]).
%
%   kappa - either the reference's anonymous curvature function (then the track must have
%           been registered once with fsae_mpc_b200_track) or a struct with fields
%           x_spline, y_spline, dl (registered on first use).
%   QP    - kept for call compatibility (the reference threads a qpOASES handle through
%           it, ltvmpc_kinetmatic_curvilinear.m:44-50); returned unchanged.

    h = fsae_mpc_b200_handle();
    if isstruct(kappa)
        fsae_mpc_b200_mex('set_track', h, 0, kappa.x_spline, kappa.y_spline, kappa.dl);
    end
    [u_opt, x_opt, exitflag, fval, slack_opt] = fsae_mpc_b200_mex('ltvmpc', h, 0, x0, x_ref, dt, x_lin, u_lin);
    exitflag = double(exitflag);
    if any(exitflag)
        display(exitflag)      % ltvmpc_kinetmatic_curvilinear.m:53-55
    end
end
