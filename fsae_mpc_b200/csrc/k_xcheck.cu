// Cross-check build only (-DFSAE_XCHECK): the shared-memory operator kernel (fused_v1.cuh), the same algorithm
// with M in shared memory (horizon 80: in an L2-resident global slab).  Not part of the product library.
#ifdef FSAE_XCHECK
#include "fused_v1.cuh"
#include "launch.h"
namespace fsae {
template <int N, bool MG>
static cudaError_t launch_v1_t(const BatchArgs& a, cudaStream_t st) {
    using S_t = SmemV1<KinModel, N, 256, MG>;
    auto kern = ltvmpc_fused_v1_kernel<KinModel, N, 256, MG>;
    static bool configured[64] = {false};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (!configured[dev & 63]) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(S_t));
        if (e != cudaSuccess) return e;
        configured[dev & 63] = true;
    }
    kern<<<a.B, 256, sizeof(S_t), st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_v1_kin(int N, const BatchArgs& a, cudaStream_t st) {
    if (N == 40) return launch_v1_t<40, false>(a, st);
    if (N == 20) return launch_v1_t<20, false>(a, st);
    if (N == 80) return launch_v1_t<80, true>(a, st);
    return cudaErrorInvalidValue;
}
size_t slab_v1_kin80() { return (size_t)Dims<KinModel, 80>::nV * Dims<KinModel, 80>::LD; }
}  // namespace fsae
#endif
