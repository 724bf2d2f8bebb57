// Roofline denominator for an FP64-FMA-bound kernel: MEASURED_PEAKS.json only carries HBM
// copy bandwidth and bf16 tensor throughput, so the DFMA peak of the part is measured here
// (dependent-chain-free DFMA streams, every SM busy), the same way the driver measures its
// own peaks: best of a few launches, CUDA events.
#pragma once
#include <cuda_runtime.h>

namespace fsae {

template <int ILP>
__global__ void __launch_bounds__(256) dfma_probe_kernel(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = (double)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true; keeps the loop
}

}  // namespace fsae
