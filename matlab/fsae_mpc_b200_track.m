function fsae_mpc_b200_track(track_id, x_spline, y_spline, dl)
%FSAE_MPC_B200_TRACK Register the arclength spline of main.m:15-17 as track `track_id`.
%Replaces building kappa = @(s) interpolate_curvature(s, x_spline, y_spline, dl) (main.m:18):
%the curvature lookup runs inside the CUDA kernel.
    fsae_mpc_b200_mex('set_track', fsae_mpc_b200_handle(), track_id, x_spline, y_spline, dl);
end
