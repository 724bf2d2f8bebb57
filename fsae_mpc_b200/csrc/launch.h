// Internal interface between the host glue (capi.cu) and the translation units that instantiate the
// fused LTV-MPC kernels (one TU per model x horizon, compiled in parallel by fsae_mpc_b200/build.py).
// Every launcher enqueues ONE kernel on `st` and returns the launch status.  `variant` selects tuning
// variants of the register-tiled kernel; only variant 2 (the product kernel) exists unless the library is
// built with -DFSAE_XCHECK (the cross-check build the tests load for fsae_debug_set_kernel_version).
#pragma once
#include <cuda_runtime.h>
#include "batch.cuh"

namespace fsae {

cudaError_t launch_kin40(const BatchArgs& a, cudaStream_t st, int variant);
cudaError_t launch_kin20(const BatchArgs& a, cudaStream_t st, int variant);
cudaError_t launch_kin80(const BatchArgs& a, cudaStream_t st, int variant);   // needs a.m_scratch: slab_kin80() doubles per problem
cudaError_t launch_dyn40(const BatchArgs& a, cudaStream_t st, int variant);
cudaError_t launch_dyn20(const BatchArgs& a, cudaStream_t st, int variant);
cudaError_t launch_dyn80(const BatchArgs& a, cudaStream_t st, int variant);   // needs a.m_scratch: slab_dyn80() doubles per problem
size_t slab_kin80();
size_t slab_dyn80();
#ifdef FSAE_XCHECK
// shared-memory operator kernel (fused_v1.cuh), kinematic only; N = 80 keeps the operator in a global slab
cudaError_t launch_v1_kin(int N, const BatchArgs& a, cudaStream_t st);
size_t slab_v1_kin80();
#endif

}  // namespace fsae
