// Staged kernels: the individual reference functions as batched device entry points.
// They exist (1) as drop-ins for callers that want one stage, and (2) as the parity taps
// the tests compare against the oracle stage by stage.  Not the hot path.
#pragma once
#include "cons.cuh"

namespace fsae {

// spline/interpolate_curvature.m:1-20, batched over s
__global__ void curvature_kernel(DevTrack tr, const double* __restrict__ s, long long n,
                                 double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = curvature(tr, s[i]);
}

// {euler,rk2,rk4}_{kinematic,dynamic}_curvilinear.m: one thread per (problem, step).
// Outputs in MATLAB layout: A [NX x NX x N x B] column-major, Bm [NX x NU x N x B], d [NX x N x B].
template <class Model>
__global__ void linearise_kernel(int B, int N, double dt, const int32_t* track_id,
                                 const int32_t* param_id, const DevTrack* tracks,
                                 const fsae_params* params, const double* __restrict__ x_lin,
                                 const double* __restrict__ u_lin, double* __restrict__ Aout,
                                 double* __restrict__ Bout, double* __restrict__ dout) {
    constexpr int NX = Model::NX, NU = Model::NU;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)B * N) return;
    const int b = (int)(t / N);
    const fsae_params& P = params[param_id ? param_id[b] : 0];
    const DevTrack tr = tracks[track_id ? track_id[b] : 0];
    double x[NX], u[NU], A[NX * NX], Bm[NX * NU], d[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = x_lin[t * NX + i];
#pragma unroll
    for (int i = 0; i < NU; ++i) u[i] = u_lin[t * NU + i];
    linearise_step<Model>(P.lin_scheme, x, u, dt, tr, P, A, Bm, d);
#pragma unroll
    for (int r = 0; r < NX; ++r)
#pragma unroll
        for (int c = 0; c < NX; ++c) Aout[t * NX * NX + c * NX + r] = A[r * NX + c];
#pragma unroll
    for (int r = 0; r < NX; ++r)
#pragma unroll
        for (int c = 0; c < NU; ++c) Bout[t * NX * NU + c * NX + r] = Bm[r * NU + c];
#pragma unroll
    for (int r = 0; r < NX; ++r) dout[t * NX + r] = d[r];
}

struct CondenseArgs {
    int B, N;
    double dt;
    const int32_t* track_id;
    const int32_t* param_id;
    const DevTrack* tracks;
    const fsae_params* params;
    const double* x0;
    const double* x_ref;
    const double* x_lin;
    const double* u_lin;
    double* H;        // nV x nV
    double* f;        // nV
    double* xA;       // nC x nV
    double* lbA;
    double* ubA;
    double* lb;
    double* ub;
    double* A_bar;    // NX*N x NX      (always a valid device buffer)
    double* B_bar;    // NX*N x nV      (always a valid device buffer)
    double* d_bar;    // NX*N           (always a valid device buffer)
    double* cconst;   // 1
};

// sequential_integration.m + *_state_constraints.m + generate_qp.m with the reference's
// dense shapes, one CTA per problem, results in global memory.  Deliberately written
// against the dense definitions (not the packed/structured forms of the fused kernel) so
// the two implementations check each other through the oracle.
template <class Model, int MAXN>
__global__ void __launch_bounds__(256) condense_kernel(CondenseArgs a) {
    using C = Cons<Model>;
    constexpr int NX = Model::NX, NU = Model::NU, NS = Model::NS;
    const int N = a.N, b = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
    const int nU = NU * N, nV = nU + NS, nXN = NX * N, nC = C::n_ref_rows(N);
    extern __shared__ __align__(16) double Ad[];           // [N][NX][NX] (dynamic: N * NX * NX * 8 bytes at launch)
    __shared__ double dd[MAXN * NX];
    __shared__ double B1[NX * NU];
    __shared__ double xf[MAXN * NX];
    __shared__ double pc[MAXN * (C::NPC > 0 ? C::NPC : 1)];
    __shared__ double g0[MAXN * C::NG0];
    __shared__ double cg[C::NCG];
    __shared__ double red[256];
    const fsae_params& P = a.params[a.param_id ? a.param_id[b] : 0];
    const DevTrack tr = a.tracks[a.track_id ? a.track_id[b] : 0];
    const double dt = a.dt;
    const double* xl = a.x_lin + (size_t)b * NX * N;
    const double* ul = a.u_lin + (size_t)b * NU * N;
    const double* x0 = a.x0 + (size_t)b * NX;
    const double* xr = a.x_ref + (size_t)b * NX * N;
    double* Abar = a.A_bar + (size_t)b * nXN * NX;
    double* Bbar = a.B_bar + (size_t)b * nXN * nV;
    double* dbar = a.d_bar + (size_t)b * nXN;

    if (tid < N) {
        const int k = tid;
        double x[NX], u[NU], Ac[NX * NX], Bc[NX * NU], dc[NX];
        for (int i = 0; i < NX; ++i) x[i] = xl[k * NX + i];
        for (int i = 0; i < NU; ++i) u[i] = ul[k * NU + i];
        linearise_step<Model>(P.lin_scheme, x, u, dt, tr, P, Ac, Bc, dc);
        for (int r = 0; r < NX; ++r)
            for (int c = 0; c < NX; ++c) Ad[(k * NX + r) * NX + c] = Ac[r * NX + c] * dt + (r == c ? 1.0 : 0.0);
        for (int r = 0; r < NX; ++r) dd[k * NX + r] = dc[r] * dt;
        if (k == 0)
            for (int i = 0; i < NX * NU; ++i) B1[i] = Bc[i] * dt;
        C::step_coefs(x, u, tr, P, pc + k * C::NPC, g0 + k * C::NG0);
    } else if (tid == N) {
        C::problem_consts(P, cg);
    }
    for (int t = tid; t < nXN * nV; t += NT) Bbar[t] = 0.0;
    __syncthreads();
    // A_bar columns (sequential_integration.m:21-26)
    if (tid < NX) {
        double v[NX], vn[NX];
        for (int r = 0; r < NX; ++r) v[r] = (r == tid) ? 1.0 : 0.0;
        for (int k = 0; k < N; ++k) {
            for (int r = 0; r < NX; ++r) {
                double acc = 0.0;
                for (int c = 0; c < NX; ++c) acc += Ad[(k * NX + r) * NX + c] * v[c];
                vn[r] = acc;
            }
            for (int r = 0; r < NX; ++r) { v[r] = vn[r]; Abar[(size_t)tid * nXN + k * NX + r] = v[r]; }
        }
    }
    // d_bar = D d(:) (sequential_integration.m:38-47) and the free response with x0
    if (tid == 32) {
        double v[NX], vn[NX], w[NX];
        for (int r = 0; r < NX; ++r) { v[r] = 0.0; w[r] = x0[r]; }
        for (int k = 0; k < N; ++k) {
            for (int r = 0; r < NX; ++r) {
                double acc = 0.0, acc2 = 0.0;
                for (int c = 0; c < NX; ++c) {
                    acc += Ad[(k * NX + r) * NX + c] * v[c];
                    acc2 += Ad[(k * NX + r) * NX + c] * w[c];
                }
                vn[r] = acc + dd[k * NX + r];
                xf[k * NX + r] = acc2 + dd[k * NX + r];
            }
            for (int r = 0; r < NX; ++r) { v[r] = vn[r]; w[r] = xf[k * NX + r]; dbar[k * NX + r] = v[r]; }
        }
    }
    // B_bar block columns (sequential_integration.m:28-36), QUIRK: B(:,:,1) on every diagonal block
    for (int t = tid; t < nU; t += NT) {
        const int i = t / NU, c = t - i * NU;
        double v[NX], vn[NX];
        for (int r = 0; r < NX; ++r) v[r] = B1[r * NU + c];
        for (int k = i; k < N; ++k) {
            if (k > i) {
                for (int r = 0; r < NX; ++r) {
                    double acc = 0.0;
                    for (int cc = 0; cc < NX; ++cc) acc += Ad[(k * NX + r) * NX + cc] * v[cc];
                    vn[r] = acc;
                }
                for (int r = 0; r < NX; ++r) v[r] = vn[r];
            }
            for (int r = 0; r < NX; ++r) Bbar[(size_t)t * nXN + k * NX + r] = v[r];
        }
    }
    __threadfence_block();
    __syncthreads();

    // generate_qp.m:23-33
    if (a.H) {
        double* H = a.H + (size_t)b * nV * nV;
        for (int t = tid; t < nV * nV; t += NT) {
            const int i = t % nV, j = t / nV;
            double acc = 0.0;
            for (int row = 0; row < nXN; ++row) {
                const int k = row / NX, r = row - k * NX;
                const double q = (k == N - 1) ? P.Q_terminal[r] : P.Q[r];
                acc += q * Bbar[(size_t)i * nXN + row] * Bbar[(size_t)j * nXN + row];
            }
            if (i == j && i < nU) acc += P.R[i % NU];
            H[t] = 2.0 * acc;
        }
    }
    if (a.f) {
        double* f = a.f + (size_t)b * nV;
        for (int j = tid; j < nV; j += NT) {
            double acc = 0.0;
            for (int row = 0; row < nXN; ++row) {
                const int k = row / NX, r = row - k * NX;
                const double q = (k == N - 1) ? P.Q_terminal[r] : P.Q[r];
                acc += q * Bbar[(size_t)j * nXN + row] * (xf[row] - xr[row]);
            }
            f[j] = (j < nU) ? 2.0 * acc : P.R_soft[j - nU];
        }
    }
    if (a.cconst) {
        double acc = 0.0;
        for (int row = tid; row < nXN; row += NT) {
            const int k = row / NX, r = row - k * NX;
            const double q = (k == N - 1) ? P.Q_terminal[r] : P.Q[r];
            const double e = xf[row] - xr[row];
            acc += q * e * e;
        }
        red[tid] = acc;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int i = 0; i < NT; ++i) s += red[i];
            a.cconst[b] = s;
        }
    }
    // ltvmpc_*_curvilinear.m:28-29
    if (a.lb && a.ub) {
        for (int j = tid; j < nV; j += NT) {
            a.lb[(size_t)b * nV + j] = (j < nU) ? P.u_lb[j % NU] : 0.0;
            a.ub[(size_t)b * nV + j] = (j < nU) ? P.u_ub[j % NU] : INFINITY;
        }
    }
    // *_state_constraints.m: xA, lbA, ubA in the reference's row order
    if (a.xA) {
        double* xA = a.xA + (size_t)b * nC * nV;
        for (int t = tid; t < nC * nV; t += NT) {
            const int row = t % nC, j = t / nC;
            int r, k, kind;
            C::ref_decode(row, N, r, k, kind);
            double v = 0.0;
            if (j < nU) {
                for (int c = 0; c < C::NXS; ++c) {
                    const double cf = C::row_coef(r, c, pc + k * C::NPC, cg);
                    if (cf != 0.0) v += cf * Bbar[(size_t)j * nXN + k * NX + C::xs_state(c)];
                }
                if (j / NU == k) v += C::row_ucoef(r, j % NU, pc + k * C::NPC, cg);
            } else {
                const int sl = C::row_slack(r);
                if (sl >= 0 && j == nU + sl) v = C::ref_slack_sign(row, N);
            }
            xA[t] = v;
        }
    }
    if (a.lbA && a.ubA) {
        for (int row = tid; row < nC; row += NT) {
            int r, k, kind;
            C::ref_decode(row, N, r, k, kind);
            double lo, up;
            C::row_bounds(r, xf + k * NX, xl + k * NX, ul + k * NU, pc + k * C::NPC, g0 + k * C::NG0, cg, P, lo, up);
            if (kind == 1) up = P.soft_far;
            else if (kind == 2) lo = -P.soft_far;
            else if (kind == 3) up = INFINITY;
            else if (kind == 4) lo = -INFINITY;
            a.lbA[(size_t)b * nC + row] = lo;
            a.ubA[(size_t)b * nC + row] = up;
        }
    }
}

}  // namespace fsae
