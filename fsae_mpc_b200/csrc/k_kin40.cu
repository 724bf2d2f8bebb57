// Kinematic model, horizon 40 (BASELINE.json configs[1]): 6 warps, 2 CTAs/SM, one constraint per search.
#include "launch_impl.cuh"
namespace fsae {
cudaError_t launch_kin40(const BatchArgs& a, cudaStream_t st, int variant) {
#ifdef FSAE_XCHECK
    switch (variant) {      // warp count x block size variants (tests, tuning)
        case 21: return launch_v2_t<KinModel, 40, 1, 8, 1, -1, false>(a, st);     // 8 warps (1 CTA/SM: shared memory)
        case 22: return launch_v2_t<KinModel, 40, 1, 12, 1, -1, false>(a, st);    // 12 warps (latency probe)
        case 23: return launch_v2_t<KinModel, 40, 1, 16, 1, -1, false>(a, st);    // 16 warps (latency probe)
        case 24: return launch_v2_t<KinModel, 40, 1, 8, 2, -1, false>(a, st);     // 8 warps, blocks of 2 (latency probe)
        case 25: return launch_v2_t<KinModel, 40, 1, 8, 3, -1, false>(a, st);     // 8 warps, blocks of 3 (latency probe)
        case 26: return launch_v2_t<KinModel, 40, 2, 6, 2, -1, false>(a, st);     // 6 warps, blocks of 2 constraints per search
        case 28: return launch_v2_t<KinModel, 40, 2, 6, 3, -1, false>(a, st);     // 6 warps, blocks of 3
        case 29: return launch_v2_t<KinModel, 40, 2, 4, 1, -1, false>(a, st);     // 4 warps
        case 31: return launch_v2_t<KinModelW, 40, 2, 6, 1, -1, false>(a, st);    // integrator coordinates, column-lane core
        case 32: return launch_v2_t<KinModelR, 40, 2, 6, 1, -1, false>(a, st);    // integrator coordinates, row-lane core
        default: break;
    }
#endif
    (void)variant;
    // Latency mode: a batch that does not fill the GPU (at most one problem per SM -- the single-vehicle loop of
    // main.m is B = 1) runs the 8-warp instantiation, whose critical path per iteration is shorter (measured on
    // B200: 174 vs 193 us for one problem, 500 vs 584 us for 148; profiles/r02_latency_variants.txt).  Larger
    // batches want the most problems in flight: 6 warps, 2 CTAs per SM.
    if (a.B <= sm_count()) return launch_v2<KinModel, 40, 1, 8, 1>(a, st);
    return launch_v2<KinModel, 40, 2, 6, 1>(a, st);
}
}  // namespace fsae
