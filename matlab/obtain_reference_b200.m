function x_ref = obtain_reference_b200(x, ds, N_s, t, s0, dt, N_t)
%OBTAIN_REFERENCE_B200  Drop-in for util/obtain_reference.m (call site main.m:115) on the GPU.
%   Same arguments and result as obtain_reference(x, ds, N_s, t, s0, dt, N_t); s0 may be a vector
%   [1 x B] (one start arclength per vehicle), then x_ref is [7 x N_t x B].
    h = fsae_mpc_b200_handle();
    x_ref = fsae_mpc_b200_mex('obtain_reference', h, x, ds, N_s, t, s0, dt, N_t);
end
