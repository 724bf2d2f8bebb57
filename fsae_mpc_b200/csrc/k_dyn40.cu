// Dynamic model (the reference's default, main.m:26), horizon 40 (BASELINE.json configs[2]).
#include "launch_impl.cuh"
namespace fsae {
cudaError_t launch_dyn40(const BatchArgs& a, cudaStream_t st, int variant) {
    (void)variant;
    return launch_v2<DynModel, 40, 1>(a, st);
}
}  // namespace fsae
