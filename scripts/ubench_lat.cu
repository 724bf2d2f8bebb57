// Latency micro-benchmarks on the target GPU (one warp): dependent DFMA, DADD, SHFL(64-bit), LDS, rcp, rsqrt, div.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double a, double b) {
    __shared__ double sm[1024];
    const int lane = threadIdx.x;
    for (int i = lane; i < 1024; i += 32) sm[i] = (double)((i * 7 + 1) % 1024);
    __syncthreads();
    double x = a + lane;
    long long t0, t1;
    const int N = 2048;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = fma(x, b, a);
    t1 = clock64(); if (lane == 0) cyc[0] = (t1 - t0);
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = x + b;
    t1 = clock64(); if (lane == 0) cyc[1] = (t1 - t0);
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1) + 0.0 * x;
    t1 = clock64(); if (lane == 0) cyc[2] = (t1 - t0);
    int idx = lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) idx = (int)sm[idx & 1023];
    t1 = clock64(); if (lane == 0) cyc[3] = (t1 - t0);
    x += idx;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = __drcp_rn(x) + 1.5;
    t1 = clock64(); if (lane == 0) cyc[4] = (t1 - t0);
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = rsqrt(x) + 1.5;
    t1 = clock64(); if (lane == 0) cyc[5] = (t1 - t0);
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = 3.0 / x + 1.5;
    t1 = clock64(); if (lane == 0) cyc[6] = (t1 - t0);
    // independent DFMA throughput for one warp (8 chains)
    double y[8];
    for (int j = 0; j < 8; ++j) y[j] = x + j;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = fma(y[j], b, a);
    }
    t1 = clock64(); if (lane == 0) cyc[7] = (t1 - t0);
    for (int j = 0; j < 8; ++j) x += y[j];
    t0 = clock64();
    for (int i = 0; i < 256; ++i) __syncthreads();
    t1 = clock64(); if (lane == 0) cyc[8] = (t1 - t0);
    // float select+add chain (ALU) for reference
    float f = (float)x;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) f = fmaf(f, 1.0001f, 0.5f);
    t1 = clock64(); if (lane == 0) cyc[9] = (t1 - t0);
    out[lane] = x + f;
}
int main() {
    double* o; long long* c; cudaMalloc(&o, 32 * 8); cudaMalloc(&c, 16 * 8);
    k<<<1, 32>>>(o, c, 1.0000001, 0.9999999); cudaDeviceSynchronize();
    k<<<1, 32>>>(o, c, 1.0000001, 0.9999999); cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
    const char* n[] = {"DFMA dep", "DADD dep", "SHFL64+DFMA dep", "LDS+cvt dep", "drcp+add dep", "rsqrt+add dep", "div+add dep", "DFMA x8 indep (per 8)", "syncthreads(1 warp)", "FFMA dep"};
    const int d[] = {2048, 2048, 2048, 2048, 2048, 2048, 2048, 2048, 256, 2048};
    for (int i = 0; i < 10; ++i) printf("%-26s %8.1f cycles\n", n[i], (double)h[i] / d[i]);
    return 0;
}
