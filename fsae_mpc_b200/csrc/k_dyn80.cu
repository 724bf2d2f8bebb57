// Dynamic model, horizon 80 (nV = 164): 12 warps, 4 of the 6 column slots of the operator tile in registers,
// 2 in shared memory; ALL B_bar rows (the four the constraints touch are 4 x 52 KB), the full H and the J
// staging in a per-problem global slab that stays L2-resident (SmemV2::BFG).
#include "launch_impl.cuh"
namespace fsae {
cudaError_t launch_dyn80(const BatchArgs& a, cudaStream_t st, int variant) {
    (void)variant;
    return launch_v2<DynModel, 80, 1, 12, 1, 4>(a, st);
}
size_t slab_dyn80() { return SmemV2<DynModel, 80, 12, 1, 4>::SLAB; }
}  // namespace fsae
