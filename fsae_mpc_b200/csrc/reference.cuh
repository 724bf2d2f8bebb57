// util/obtain_reference.m for B vehicles on one planned lap: thread b re-parameterises the plan (state
// samples every ds metres of arclength, time t spent in each segment) from arclength to time, starting
// at its own arclength s0.  The time-stepping loop of obtain_reference.m:24-35 is inherently sequential
// per vehicle (a few segments per horizon step), so the batch is the parallel dimension.  Every floating
// point operation is written with round-to-nearest intrinsics in the reference's order (no FMA
// contraction): the result is bit-identical to the .m file's.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace fsae {

// MATLAB mod(a, m), m > 0 (fmod is exact)
__device__ __forceinline__ double matlab_mod(double a, double m) {
    double r = fmod(a, m);
    if (r != 0.0 && ((r < 0.0) != (m < 0.0))) r += m;
    return r;
}

__global__ void obtain_reference_kernel(const double* __restrict__ x, const double* __restrict__ t, int N_s, double ds,
                                        const double* __restrict__ s0v, int B, double dt, int N_t,
                                        double* __restrict__ x_ref) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double s0 = s0v[b], L = __dmul_rn(ds, (double)N_s);                        // :5
    const double q0 = matlab_mod(s0, L) / ds;
    const int idx0 = (int)floor(q0) + 1;                                             // :21 (one-based)
    const double rto0 = matlab_mod(q0, 1.0);                                         // :22
    int idx = idx0;
    double rto = rto0;
    double* out = x_ref + (size_t)b * 7 * N_t;
    for (int i = 1; i <= N_t; ++i) {                                                 // :24-35
        const int idxp = idx;
        const double rtop = rto;
        double t_rem = dt;
        rto = __dadd_rn(rtop, t_rem / t[idx - 1]);
        t_rem = __dsub_rn(t_rem, __dmul_rn(t[idxp - 1], __dsub_rn(1.0, rtop)));
        for (int guard = 0; rto > 1.0 && guard <= N_s; ++guard) {                   // a malformed plan cannot hang the GPU
            idx = idx % N_s + 1;                                                     // nxt(), :58-60
            rto = t_rem / t[idx - 1];
            t_rem = __dsub_rn(t_rem, t[idx - 1]);
        }
        const double w = __dsub_rn(__dsub_rn(__dadd_rn((double)idx, rto), (double)idx0), rto0);
        out[(i - 1) * 7] = __dadd_rn(s0, __dmul_rn(matlab_mod(w, (double)N_s), ds));   // :41
        const int a = idx - 1, bn = idx % N_s;
#pragma unroll
        for (int k = 0; k < 6; ++k) {                                                // :42-47  n, mu, x_d, y_d, theta_d, delta
            const double va = x[a * 8 + k], vb = x[bn * 8 + k];
            out[(i - 1) * 7 + 1 + k] = __dadd_rn(va, __dmul_rn(__dsub_rn(vb, va), rto));
        }
    }
}

}  // namespace fsae
