// Kinematic model, horizon 20: 4 warps x 5 CTAs/SM (13 KB operator).
#include "launch_impl.cuh"
namespace fsae {
cudaError_t launch_kin20(const BatchArgs& a, cudaStream_t st, int variant) {
#ifdef FSAE_XCHECK
    switch (variant) {
        case 21: return launch_v2_t<KinModel, 20, 2, 8, 1, -1, false>(a, st);     // 8 warps x 2 CTAs/SM: 3.36M QP/s
        case 26: return launch_v2_t<KinModel, 20, 3, 6, 1, -1, false>(a, st);     // 6 warps x 3: 4.39M
        case 31: return launch_v2_t<KinModelW, 20, 5, 4, 1, -1, false>(a, st);    // integrator coordinates, column-lane core
        case 32: return launch_v2_t<KinModelR, 20, 5, 4, 1, -1, false>(a, st);    // integrator coordinates, row-lane core
        default: break;
    }
#endif
    (void)variant;
    if (a.B <= sm_count()) return launch_v2<KinModel, 20, 2, 8, 1>(a, st);     // latency mode (see k_kin40.cu): 86 vs 112 us
    return launch_v2<KinModel, 20, 5, 4, 1>(a, st);                  // 4 warps x 5 CTAs/SM
}
}  // namespace fsae
