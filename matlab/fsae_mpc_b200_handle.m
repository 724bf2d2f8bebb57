function h = fsae_mpc_b200_handle(device)
%FSAE_MPC_B200_HANDLE Process-wide context handle of the CUDA library (created on first use).
    persistent H
    if isempty(H)
        if nargin < 1, device = 0; end
        H = fsae_mpc_b200_mex('create', device);
    end
    h = H;
end
