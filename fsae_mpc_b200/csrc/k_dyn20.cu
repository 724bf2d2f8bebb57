// Dynamic model, horizon 20.
#include "launch_impl.cuh"
namespace fsae {
cudaError_t launch_dyn20(const BatchArgs& a, cudaStream_t st, int variant) {
    (void)variant;
    return launch_v2<DynModel, 20, 1>(a, st);
}
}  // namespace fsae
