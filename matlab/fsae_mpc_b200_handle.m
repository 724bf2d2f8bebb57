function h = fsae_mpc_b200_handle(devices)
%FSAE_MPC_B200_HANDLE Process-wide handle of the CUDA library (created on first use).
%The handle is a DEVICE POOL: with no argument it takes every visible B200, so that one MATLAB process drives
%the 8 GPUs of a box and batched ltvmpc_*_curvilinear_b200 calls are split over them (fsae_ltvmpc_host_pool);
%fsae_mpc_b200_handle(0) or fsae_mpc_b200_handle([0 1]) restricts it (first call only).
    persistent H
    if isempty(H)
        if nargin < 1, devices = []; end
        H = fsae_mpc_b200_mex('create', devices);
    end
    h = H;
end
