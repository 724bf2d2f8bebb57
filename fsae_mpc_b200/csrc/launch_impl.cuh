// Launch helper shared by the kernel translation units.
#pragma once
#include "fused_v2.cuh"
#include "launch.h"

namespace fsae {

template <class Model, int N, int MINB, int NW = 8, int KB = 1, int CSR = -1, bool PAD = false>
static cudaError_t launch_v2_t(const BatchArgs& a, cudaStream_t st) {
    using S_t = SmemV2<Model, N, NW, KB, CSR>;
    // the occupancy the kernel was tuned for must survive every change of the shared-memory layout:
    // 228 KB per SM, 1 KB reserved per resident CTA
    static_assert((size_t)MINB * (sizeof(S_t) + 1024) <= 233472, "MINB CTAs per SM no longer fit shared memory");
    auto kern = ltvmpc_fused_v2_kernel<Model, N, MINB, NW, KB, CSR, PAD>;
    static bool configured[64] = {false};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (!configured[dev & 63]) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(S_t));
        if (e != cudaSuccess) return e;
        configured[dev & 63] = true;
    }
    kern<<<a.B, 32 * NW, sizeof(S_t), st>>>(a);
    return cudaGetLastError();
}

// SMs of the current device (latency-mode dispatch)
static int sm_count() {
    static int n[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (!n[dev & 63]) cudaDeviceGetAttribute(&n[dev & 63], cudaDevAttrMultiProcessorCount, dev);
    return n[dev & 63];
}

// a.N == N: the exact-capacity kernel; a.N < N: the padding kernel (runtime horizon)
template <class Model, int N, int MINB, int NW = 8, int KB = 1, int CSR = -1>
static cudaError_t launch_v2(const BatchArgs& a, cudaStream_t st) {
    if (a.N == N) return launch_v2_t<Model, N, MINB, NW, KB, CSR, false>(a, st);
    if (a.N < 1 || a.N > N) return cudaErrorInvalidValue;
    return launch_v2_t<Model, N, MINB, NW, KB, CSR, true>(a, st);
}

}  // namespace fsae
