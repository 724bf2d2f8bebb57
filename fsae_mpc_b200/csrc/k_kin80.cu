// Kinematic model, horizon 80 (nV = 161): 12 warps, 4 of the 6 column slots of the operator tile in registers,
// 2 in shared memory; B_bar rows the constraints do not touch, the full H and the J staging in a per-problem
// global slab that stays L2-resident.
#include "launch_impl.cuh"
namespace fsae {
cudaError_t launch_kin80(const BatchArgs& a, cudaStream_t st, int variant) {
    (void)variant;
    return launch_v2<KinModel, 80, 1, 12, 1, 4>(a, st);
}
size_t slab_kin80() { return SmemV2<KinModel, 80, 12, 1, 4>::SLAB; }
}  // namespace fsae
