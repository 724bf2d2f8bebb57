// Fused per-problem LTV-MPC step, version 1: one CTA per problem, the dual active-set
// operator M = [K1 | J2] in SHARED memory.  Kept as the simple, obviously-correct variant
// (and as the parity cross-check of the register-tiled v2 kernel).
//
// Pipeline per CTA (reference function each stage replaces):
//   load        x0, x_ref, x_lin, u_lin                      (arguments of ltvmpc_*_curvilinear.m:1)
//   linearise   A_k, B_k, d_k per step                        (rk2_kinematic_curvilinear.m:25-50 ...)
//   discretise  A_k*dt+I, B*dt, d*dt, free response           (sequential_integration.m:16-18,21-26,38-47)
//   condense    packed B_bar rows, H, g, row bounds           (sequential_integration.m:28-36,
//                                                              *_state_constraints.m, generate_qp.m:23-33)
//   factor      H = L L',  J = L^-T                           (inside qpOASES in the reference)
//   solve       Goldfarb-Idnani dual active set, K-form       (qpOASES call, ltvmpc_*_curvilinear.m:52)
//   output      u_opt, slack_opt, x_opt, fval, exitflag       (ltvmpc_*_curvilinear.m:57-60)
#pragma once
#include "batch.cuh"

namespace fsae {

template <class Model, int N, int NT, bool MG = false>
struct SmemV1 {
    using D = Dims<Model, N>;
    using C = Cons<Model>;
    static constexpr int NW = NT / 32;
    double M[MG ? 1 : D::nV * D::LD];   // MG: the operator lives in a per-CTA global (L2) slab instead
    double Hp[D::HP];
    double Bc[C::NCR * D::NPK];
    double Ad[N * C::NREAL * D::NX];
    double B1[D::NX * D::NU];
    double xf[N * D::NX];
    double xl[N * D::NX];
    double ul[N * D::NU];
    double pc[N * C::NPC];
    double g0[N * C::NG0];
    double cg[C::NCG];
    double rlo[D::NROWS], rup[D::NROWS];
    double x[D::nV], g[D::nV], y[D::nV], z[D::nV], wv[D::nV], kv[D::nV], nv[D::nV], lam[D::nV];
    double part[3 * D::nV];
    double xs[C::NXS * N];
    double dvec[D::nV];
    double red_val[NW];
    double sc[16];           // scalars: 0 t, 1 delta2, 2 sgn*delta, 3 beta, 4 kHk, 5 lam_p, 6 s_p, 7 cost const
    int red_idx[NW];
    int act[D::nV];          // slot*2 + (side>0)
    int isc[16];             // 0 q, 1 p slot, 2 p side, 3 step type, 4 l, 5 iters, 6 exit, 7 refreshes
    int8_t status[D::NSLOT + 8];
};

enum { STEP_FULL = 0, STEP_PARTIAL = 1, STEP_DUAL = 2, STEP_INFEAS = 3 };

template <int NT>
__device__ __forceinline__ void block_argmin(double v, int idx, double* red_val, int* red_idx,
                                             double& out_v, int& out_i) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    if (lane == 0) { red_val[w] = v; red_idx[w] = idx; }
    __syncthreads();
    out_v = red_val[0];
    out_i = red_idx[0];
#pragma unroll
    for (int k = 1; k < NT / 32; ++k) {
        const double ov = red_val[k];
        const int oi = red_idx[k];
        if (ov < out_v || (ov == out_v && oi < out_i)) { out_v = ov; out_i = oi; }
    }
    __syncthreads();   // red_* reusable afterwards
}


// operator element load: shared memory, or (global-operator variant) an L1-bypassing global load
// so that values written by other warps of the CTA before a barrier are always observed
template <bool MG>
__device__ __forceinline__ double ldM(const double* p) {
    if constexpr (MG) return __ldcg(p);
    else return *p;
}

// y = M^T v over all columns: thread (j, part) partial sums; caller syncs then sums 3 parts.
template <class D, int NT, bool MG = false>
__device__ __forceinline__ void matvec_T_parts(const double* M, const double* v, double* part) {
    constexpr int nV = D::nV, LD = D::LD;
    constexpr int CH = (nV + 2) / 3;
    for (int t = threadIdx.x; t < 3 * nV; t += NT) {
        const int pt = t / nV, j = t - pt * nV;
        const int i0 = pt * CH, i1 = (i0 + CH < nV) ? i0 + CH : nV;
        double acc = 0.0;
        for (int i = i0; i < i1; ++i) acc += ldM<MG>(&M[i * LD + j]) * v[i];
        part[pt * nV + j] = acc;
    }
}

// z = M[:, q:] * y[q:]: thread (i, part) partial sums over a third of the column range.
template <class D, int NT, bool MG = false>
__device__ __forceinline__ void matvec_N_parts(const double* M, const double* y, int q, double* part) {
    constexpr int nV = D::nV, LD = D::LD;
    const int span = nV - q;
    const int CH = (span + 2) / 3;
    for (int t = threadIdx.x; t < 3 * nV; t += NT) {
        const int pt = t / nV, i = t - pt * nV;
        const int j0 = q + pt * CH;
        const int j1 = (j0 + CH < nV) ? j0 + CH : nV;
        double acc = 0.0;
        for (int j = j0; j < j1; ++j) acc += ldM<MG>(&M[i * LD + j]) * y[j];
        part[pt * nV + i] = acc;
    }
}

// w = Hp (packed symmetric) * v, thread per row
template <class D, int NT>
__device__ __forceinline__ void symv_packed(const double* Hp, const double* v, double* out) {
    constexpr int nV = D::nV;
    for (int i = threadIdx.x; i < nV; i += NT) {
        double acc = 0.0;
        for (int j = 0; j <= i; ++j) acc += Hp[D::hp(i, j)] * v[j];
        for (int j = i + 1; j < nV; ++j) acc += Hp[D::hp(j, i)] * v[j];
        out[i] = acc;
    }
}

// constraint-state perturbations xs[c][k] = (B_bar x_u)[state c, step k]
template <class Model, int N, int NT, bool MG>
__device__ __forceinline__ void eval_xs(const SmemV1<Model, N, NT, MG>& S, const double* x, double dt,
                                        double* xs) {
    using D = Dims<Model, N>;
    using C = Cons<Model>;
    for (int t = threadIdx.x; t < C::NXS * N; t += NT) {
        const int c = t / N, k = t - c * N;
        double acc = 0.0;
        if (c < C::NCR) {
            const double* row = S.Bc + c * D::NPK + D::pk(k, 0);
            const int len = D::NU * (k + 1);
            for (int j = 0; j < len; ++j) acc += row[j] * x[j];
        } else {
            const int uc = C::int_ucol(c - C::NCR);
            for (int i = 0; i <= k; ++i) acc += x[D::NU * i + uc];
            acc *= dt;
        }
        xs[t] = acc;
    }
}

// entry j of the GI normal of (slot, side):  n'x >= b  form
template <class Model, int N, int NT, bool MG>
__device__ __forceinline__ double normal_entry(const SmemV1<Model, N, NT, MG>& S, int slot, int side,
                                               int j, double dt) {
    using D = Dims<Model, N>;
    using C = Cons<Model>;
    const double sg = side < 0 ? 1.0 : -1.0;
    if (slot < D::nV) return j == slot ? sg : 0.0;
    const int rr = slot - D::nV;
    const int r = rr / N, k = rr - r * N;
    if (j >= D::nU) {
        const int sl = C::row_slack(r);
        return (sl >= 0 && j == D::nU + sl) ? 1.0 : 0.0;
    }
    const int step = j / D::NU, uc = j - step * D::NU;
    if (step > k) return 0.0;
    const double* pc = S.pc + k * C::NPC;
    double v = 0.0;
#pragma unroll
    for (int c = 0; c < C::NCR; ++c) v += C::row_coef(r, c, pc, S.cg) * S.Bc[c * D::NPK + D::pk(k, j)];
#pragma unroll
    for (int c = 0; c < C::NINT; ++c)
        if (C::int_ucol(c) == uc) v += C::row_coef(r, C::NCR + c, pc, S.cg) * dt;
    if (step == k) v += C::row_ucoef(r, uc, pc, S.cg);
    return sg * v;
}

template <class Model, int N, int NT, bool MG = false>
__global__ void __launch_bounds__(NT, 1) ltvmpc_fused_v1_kernel(BatchArgs a) {
    using D = Dims<Model, N>;
    using C = Cons<Model>;
    using S_t = SmemV1<Model, N, NT, MG>;
    constexpr int NX = D::NX, NU = D::NU, NS = D::NS, nU = D::nU, nV = D::nV, LD = D::LD;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    S_t& S = *reinterpret_cast<S_t*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    if (b >= a.B) return;
    const fsae_params& P = a.params[a.param_id ? a.param_id[b] : 0];
    const DevTrack tr = a.tracks[a.track_id ? a.track_id[b] : 0];
    const double dt = a.dt;
    const int scheme = P.lin_scheme;
    double* const Mop = MG ? (a.m_scratch + (size_t)b * D::nV * D::LD) : S.M;

    // ---------------------------------------------------------------- load
    {
        const double* gxl = a.x_lin + (size_t)b * NX * N;
        const double* gul = a.u_lin + (size_t)b * NU * N;
        for (int i = tid; i < NX * N; i += NT) S.xl[i] = gxl[i];
        for (int i = tid; i < NU * N; i += NT) S.ul[i] = gul[i];
        for (int i = tid; i < D::NSLOT; i += NT) S.status[i] = 0;
    }
    __syncthreads();

    // ---------------------------------------------------------------- linearise + discretise
    // dd (discrete affine term) is parked in S.part/S.y/... until the free response is done.
    double* dd = S.part;            // needs N*NX <= 3*nV  (200 <= 243; 280 <= 252 NO for dynamic)
    static_assert(N * NX <= 3 * nV + nV, "dd scratch too small");
    if (tid < N) {
        const int k = tid;
        double Ac[NX * NX], Bc_[NX * NU], dc[NX];
        linearise_step<Model>(scheme, S.xl + k * NX, S.ul + k * NU, dt, tr, P, Ac, Bc_, dc);
        // sequential_integration.m:16-18
#pragma unroll
        for (int i = 0; i < C::NREAL; ++i) {
            const int r = C::real_state(i);
#pragma unroll
            for (int c = 0; c < NX; ++c)
                S.Ad[(k * C::NREAL + i) * NX + c] = Ac[r * NX + c] * dt + (r == c ? 1.0 : 0.0);
        }
#pragma unroll
        for (int r = 0; r < NX; ++r) dd[k * NX + r] = dc[r] * dt;
        if (k == 0) {
#pragma unroll
            for (int i = 0; i < NX * NU; ++i) S.B1[i] = Bc_[i] * dt;   // QUIRK: B(:,:,1) everywhere
        }
        C::step_coefs(S.xl + k * NX, S.ul + k * NU, tr, P, S.pc + k * C::NPC, S.g0 + k * C::NG0);
    } else if (tid == N) {
        C::problem_consts(P, S.cg);
    }
    __syncthreads();

    // ---------------------------------------------------------------- free response + B_bar chains
    double* Bf = Mop;               // NREAL packed rows, lives in M until H is built
    static_assert(C::NREAL * D::NPK <= nV * LD, "packed B_bar does not fit the M region");
    if (warp == NT / 32 - 1) {
        // xf_k = Ad_k xf_{k-1} + dd_k, xf_{-1} = x0   (A_bar x0 + d_bar, sequential_integration.m:21-26,38-47)
        if (lane == 0) {
            double xp[NX], xn[NX];
            const double* gx0 = a.x0 + (size_t)b * NX;
#pragma unroll
            for (int i = 0; i < NX; ++i) xp[i] = gx0[i];
            for (int k = 0; k < N; ++k) {
#pragma unroll
                for (int i = 0; i < NX; ++i) xn[i] = xp[i];       // integrator rows: identity
#pragma unroll
                for (int i = 0; i < C::NREAL; ++i) {
                    const int r = C::real_state(i);
                    double acc = 0.0;
#pragma unroll
                    for (int c = 0; c < NX; ++c) acc += S.Ad[(k * C::NREAL + i) * NX + c] * xp[c];
                    xn[r] = acc;
                }
#pragma unroll
                for (int i = 0; i < NX; ++i) {
                    xn[i] += dd[k * NX + i];
                    S.xf[k * NX + i] = xn[i];
                    xp[i] = xn[i];
                }
            }
        }
    } else {
        // column (i, c) of B_bar: v = B1[:, c] at step i, then v <- Ad_k v  (sequential_integration.m:28-36)
        for (int t = tid; t < N * NU; t += NT - 32) {
            const int i = t / NU, c = t - i * NU;
            double v[NX], vn[NX];
#pragma unroll
            for (int r = 0; r < NX; ++r) v[r] = S.B1[r * NU + c];
            for (int k = i; k < N; ++k) {
                if (k > i) {
#pragma unroll
                    for (int r = 0; r < NX; ++r) vn[r] = v[r];
#pragma unroll
                    for (int ii = 0; ii < C::NREAL; ++ii) {
                        double acc = 0.0;
#pragma unroll
                        for (int cc = 0; cc < NX; ++cc) acc += S.Ad[(k * C::NREAL + ii) * NX + cc] * v[cc];
                        vn[C::real_state(ii)] = acc;
                    }
#pragma unroll
                    for (int r = 0; r < NX; ++r) v[r] = vn[r];
                }
#pragma unroll
                for (int ii = 0; ii < C::NREAL; ++ii) Bf[ii * D::NPK + D::pk(k, t)] = v[C::real_state(ii)];
            }
        }
    }
    __syncthreads();

    // ---------------------------------------------------------------- H (packed), g, bounds
    {
        // tracking error e[k][r] = xf - x_ref -> reuse S.part after dd is dead (xf done)
        double* e = S.part;
        const double* gxr = a.x_ref + (size_t)b * NX * N;
        __syncthreads();
        for (int i = tid; i < NX * N; i += NT) e[i] = S.xf[i] - gxr[i];
        __syncthreads();
        // generate_qp.m:29: H = 2 (B' Qbar B + Rbar)
        for (int t = tid; t < D::HP; t += NT) {
            // invert packed index t -> (i >= j)
            int i = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
            while (D::hp(i + 1, 0) <= t) ++i;
            while (D::hp(i, 0) > t) --i;
            const int j = t - D::hp(i, 0);
            double acc = 0.0;
            if (i < nU) {
                const int si = i / NU, ci = i - si * NU, cj = j % NU;
                for (int k = si; k < N; ++k) {
                    double s = 0.0;
#pragma unroll
                    for (int ii = 0; ii < C::NREAL; ++ii) {
                        const int r = C::real_state(ii);
                        const double q = (k == N - 1) ? P.Q_terminal[r] : P.Q[r];
                        s += q * ldM<MG>(&Bf[ii * D::NPK + D::pk(k, i)]) * ldM<MG>(&Bf[ii * D::NPK + D::pk(k, j)]);
                    }
                    acc += s;
                }
#pragma unroll
                for (int ii = 0; ii < C::NINT; ++ii) {
                    if (C::int_ucol(ii) == ci && cj == ci) {
                        const int r = C::int_state(ii);
                        acc += dt * dt * (P.Q[r] * (double)(N - 1 - si) + P.Q_terminal[r]);
                    }
                }
                acc *= 2.0;
                if (i == j) acc += 2.0 * P.R[ci];
            } else if (i == j) {
                acc = P.flat_eps;      // slack: zero curvature in the reference (generate_qp.m:25)
            }
            S.Hp[t] = acc;
        }
        // generate_qp.m:30-31: f = 2 B' Qbar e ; f(slack) = R_soft
        for (int j = tid; j < nV; j += NT) {
            double acc = 0.0;
            if (j < nU) {
                const int sj = j / NU, cj = j - sj * NU;
                for (int k = sj; k < N; ++k) {
#pragma unroll
                    for (int ii = 0; ii < C::NREAL; ++ii) {
                        const int r = C::real_state(ii);
                        const double q = (k == N - 1) ? P.Q_terminal[r] : P.Q[r];
                        acc += q * ldM<MG>(&Bf[ii * D::NPK + D::pk(k, j)]) * e[k * NX + r];
                    }
#pragma unroll
                    for (int ii = 0; ii < C::NINT; ++ii) {
                        if (C::int_ucol(ii) == cj) {
                            const int r = C::int_state(ii);
                            const double q = (k == N - 1) ? P.Q_terminal[r] : P.Q[r];
                            acc += q * dt * e[k * NX + r];
                        }
                    }
                }
                acc *= 2.0;
            } else {
                acc = P.R_soft[j - nU];
            }
            S.g[j] = acc;
        }
        // generate_qp.m:33 const = e' Qbar e
        if (warp == 0) {
            double acc = 0.0;
            for (int i = lane; i < NX * N; i += 32) {
                const int k = i / NX, r = i - k * NX;
                const double q = (k == N - 1) ? P.Q_terminal[r] : P.Q[r];
                acc += q * e[i] * e[i];
            }
            acc = warp_sum(acc);
            if (lane == 0) S.sc[7] = acc;
        }
        // row bounds
        for (int t = tid; t < D::NROWS; t += NT) {
            const int r = t / N, k = t - r * N;
            double lo, up;
            C::row_bounds(r, S.xf + k * NX, S.xl + k * NX, S.ul + k * NU, S.pc + k * C::NPC, S.g0 + k * C::NG0, S.cg, P, lo, up);
            S.rlo[t] = lo;
            S.rup[t] = up;
        }
        // keep the constraint rows of B_bar
        for (int t = tid; t < C::NCR * D::NPK; t += NT) {
            const int c = t / D::NPK, o = t - c * D::NPK;
            S.Bc[t] = ldM<MG>(&Bf[C::cons_real(c) * D::NPK + o]);
        }
    }
    __syncthreads();
    if (a.dbg_H) {
        double* gH = a.dbg_H + (size_t)b * nV * nV;
        for (int t = tid; t < nV * nV; t += NT) {
            const int i = t % nV, j = t / nV;
            double v = (i >= j) ? S.Hp[D::hp(i, j)] : S.Hp[D::hp(j, i)];
            if (i >= nU && i == j) v = 0.0;
            gH[t] = v;
        }
        double* gg = a.dbg_g + (size_t)b * nV;
        for (int t = tid; t < nV; t += NT) gg[t] = S.g[t];
    }

    // ---------------------------------------------------------------- factor: H_uu = L1 D L1', X = L1^-1
    // Working storage: MU(i,j) = M[i*LD + NS + j], i,j < nU.  Lower triangle of the active
    // trailing block holds the Schur complement; eliminated columns hold X = L1^-1 (unit lower).
    __syncthreads();
#define MU(i, j) Mop[(i) * LD + NS + (j)]
#define MUL(i, j) ldM<MG>(&Mop[(i) * LD + NS + (j)])
    for (int t = tid; t < nU * nU; t += NT) {
        const int i = t / nU, j = t - i * nU;
        MU(i, j) = (i >= j) ? S.Hp[D::hp(i, j)] : 0.0;
    }
    __syncthreads();
    for (int k = 0; k < nU; ++k) {
        const double piv = MUL(k, k);
        const double rp = 1.0 / piv;
        for (int i = k + 1 + tid; i < nU; i += NT) {
            const double cik = MUL(i, k);
            S.z[i] = cik;                 // column k of the Schur complement
            S.kv[i] = cik * rp;           // multipliers l_i
        }
        if (tid == 0) S.dvec[k] = piv;
        __syncthreads();
        // rows i > k: trailing update (cols k < j <= i), X update (cols c < k), X[i][k] = -l_i
        const int rows = nU - 1 - k;
        for (int t = tid; t < rows * nU; t += NT) {
            const int i = k + 1 + t / nU, j = t % nU;
            const double li = S.kv[i];
            if (j < k) MU(i, j) = MUL(i, j) - li * MUL(k, j);
            else if (j == k) MU(i, k) = -li;
            else if (j <= i) MU(i, j) = MUL(i, j) - li * S.z[j];
        }
        __syncthreads();
    }
    // J = L^-T: J[r][c] = X[c][r] / sqrt(d_c)  (r <= c), zero below.  In place: upper <- lower^T.
    for (int t = tid; t < nU * nU; t += NT) {
        const int c = t / nU, r = t - c * nU;      // c >= r pairs only
        if (r < c) MU(r, c) = MUL(c, r) * rsqrt(S.dvec[c]);
    }
    __syncthreads();
    for (int t = tid; t < nU * nU; t += NT) {
        const int i = t / nU, j = t - i * nU;
        if (i == j) MU(i, i) = rsqrt(S.dvec[i]);
        else if (i > j) MU(i, j) = 0.0;
    }
    // slack rows/columns: K1 columns 0..NS-1 = e_{nU+k}; slack rows of J2 are zero
    for (int t = tid; t < nV * NS; t += NT) {
        const int i = t / NS, k = t - i * NS;
        Mop[i * LD + k] = (i == nU + k) ? 1.0 : 0.0;
    }
    for (int t = tid; t < NS * nU; t += NT) {
        const int k = t / nU, j = t - k * nU;
        Mop[(nU + k) * LD + NS + j] = 0.0;
    }
#undef MU
#undef MUL
    if (tid < NS) {
        S.act[tid] = (nU + tid) * 2;            // slack lower bound active
        S.lam[tid] = S.g[nU + tid];
        S.status[nU + tid] = -1;
    }
    if (tid == 0) {
        S.isc[0] = NS;
        S.isc[5] = 0;
        S.isc[6] = FSAE_EXIT_SOLVED;
        S.isc[7] = 0;
    }
    __syncthreads();

    // ---------------------------------------------------------------- x0 = -J2 J2' g ; slack at 0
    matvec_T_parts<D, NT, MG>(Mop, S.g, S.part);
    __syncthreads();
    for (int j = tid; j < nV; j += NT) S.y[j] = S.part[j] + S.part[nV + j] + S.part[2 * nV + j];
    __syncthreads();
    matvec_N_parts<D, NT, MG>(Mop, S.y, NS, S.part);
    __syncthreads();
    for (int i = tid; i < nV; i += NT)
        S.x[i] = (i < nU) ? -(S.part[i] + S.part[nV + i] + S.part[2 * nV + i]) : 0.0;
    __syncthreads();

    // ---------------------------------------------------------------- Goldfarb-Idnani, K-form
    const double tol = P.feas_tol;
    const int max_iter = P.max_iter > 0 ? P.max_iter : 5 * (nV + C::n_ref_rows(N));
    unsigned long long n_add = 0, n_drop = 0, drops_at_refresh = 0;
    // drop working-set column l (all threads; barrier on exit)
    auto drop_column = [&](int l) {
        const int q = S.isc[0];
        for (int i = tid; i < nV; i += NT) S.kv[i] = ldM<MG>(&Mop[i * LD + l]);
        __syncthreads();
        symv_packed<D, NT>(S.Hp, S.kv, S.wv);
        __syncthreads();
        matvec_T_parts<D, NT, MG>(Mop, S.wv, S.part);     // all columns; only j<q, j!=l used
        if (warp == 0) {
            double acc = 0.0;
            for (int i = lane; i < nV; i += 32) acc += S.kv[i] * S.wv[i];
            acc = warp_sum(acc);
            if (lane == 0) S.sc[4] = acc;
        }
        __syncthreads();
        const double kHk = S.sc[4];
        for (int j = tid; j < q; j += NT)
            S.z[j] = -(S.part[j] + S.part[nV + j] + S.part[2 * nV + j]) / kHk;   // r'
        __syncthreads();
        for (int tt = tid; tt < nV * q; tt += NT) {
            const int i = tt / q, j = tt - i * q;
            if (j != l) Mop[i * LD + j] = ldM<MG>(&Mop[i * LD + j]) + S.kv[i] * S.z[j];
        }
        __syncthreads();
        const double rs = rsqrt(kHk);
        for (int i = tid; i < nV; i += NT) {
            if (l != q - 1) Mop[i * LD + l] = ldM<MG>(&Mop[i * LD + q - 1]);
            Mop[i * LD + q - 1] = S.kv[i] * rs;
        }
        if (tid == 0) {
            S.status[S.act[l] >> 1] = 0;
            S.act[l] = S.act[q - 1];
            S.lam[l] = S.lam[q - 1];
            S.isc[0] = q - 1;
        }
++n_drop;
        __syncthreads();
    };
    while (true) {
        // P1: most violated inactive constraint side
        eval_xs<Model, N, NT, MG>(S, S.x, dt, S.xs);
        __syncthreads();
        double best = 0.0;
        int best_i = 0x7fffffff;
        for (int slot = tid; slot < D::NSLOT; slot += NT) {
            if (S.status[slot] != 0) continue;
            double vlo, vup;
            if (slot < nV) {
                const double xv = S.x[slot];
                const double lb = (slot < nU) ? P.u_lb[slot % NU] : 0.0;
                const double ub = (slot < nU) ? P.u_ub[slot % NU] : INFINITY;
                vlo = xv - lb;
                vup = ub - xv;
            } else {
                const int rr = slot - nV, r = rr / N, k = rr - r * N;
                double xsk[C::NXS];
#pragma unroll
                for (int c = 0; c < C::NXS; ++c) xsk[c] = S.xs[c * N + k];
                const double rv = C::row_value(r, xsk, S.pc + k * C::NPC, S.cg, S.x[NU * k]);
                const int sl = C::row_slack(r);
                const double sv = sl >= 0 ? S.x[nU + sl] : 0.0;
                vlo = rv + sv - S.rlo[rr];
                vup = S.rup[rr] - rv + sv;
            }
            if (vlo < best) { best = vlo; best_i = slot * 2; }
            if (vup < best) { best = vup; best_i = slot * 2 + 1; }
        }
        double viol;
        int pcode;
        block_argmin<NT>(best, best_i, S.red_val, S.red_idx, viol, pcode);
        if (!(viol < -tol)) {
            // converged on this working set: if partial steps happened since the last refresh, a
            // Newton step on the active manifold removes the round-off they leave and the multipliers
            // are recomputed from stationarity; a column whose recomputed multiplier is negative is
            // dropped and the step repeated (same rule as gi_core.cuh).
            if (n_drop == drops_at_refresh) break;
            for (int pass = 0;; ++pass) {
                const int q = S.isc[0];
                symv_packed<D, NT>(S.Hp, S.x, S.wv);
                __syncthreads();
                for (int i = tid; i < nV; i += NT) S.wv[i] += S.g[i];
                __syncthreads();
                matvec_T_parts<D, NT, MG>(Mop, S.wv, S.part);
                __syncthreads();
                for (int j = tid; j < nV; j += NT) S.y[j] = S.part[j] + S.part[nV + j] + S.part[2 * nV + j];
                __syncthreads();
                matvec_N_parts<D, NT, MG>(Mop, S.y, q, S.part);
                __syncthreads();
                for (int i = tid; i < nV; i += NT) S.x[i] -= S.part[i] + S.part[nV + i] + S.part[2 * nV + i];
                double ymin = 0.0, ymaxn = 0.0;
                int lmin = 0x7fffffff;
                for (int j = tid; j < q; j += NT) {
                    const double yj = S.y[j];
                    S.lam[j] = fmax(yj, 0.0);   // lam = K1'(Hx+g) (pre-step grad; J2 part is H-orthogonal to K1)
                    ymaxn = fmin(ymaxn, -fabs(yj));
                    if (yj < ymin) { ymin = yj; lmin = j; }
                }
                double gmin, gmaxn;
                int gl, dummy;
                block_argmin<NT>(ymin, lmin, S.red_val, S.red_idx, gmin, gl);
                __syncthreads();
                block_argmin<NT>(ymaxn, tid, S.red_val, S.red_idx, gmaxn, dummy);
                if (tid == 0) S.isc[7] += 1;
                __syncthreads();
                drops_at_refresh = n_drop;
                if (!(gmin < -1e-10 * (1.0 - gmaxn)) || pass >= 8) break;
                drop_column(gl);
                drops_at_refresh = n_drop;
            }
            continue;
        }
        const int pslot = pcode >> 1, pside = (pcode & 1) ? +1 : -1;
        // P2: normal
        for (int j = tid; j < nV; j += NT) S.nv[j] = normal_entry<Model, N, NT, MG>(S, pslot, pside, j, dt);
        if (tid == 0) S.sc[5] = 0.0;     // multiplier of p
        __syncthreads();
        bool done_p = false;
        while (!done_p) {
            const int q = S.isc[0];
            // P3: y = M' n
            matvec_T_parts<D, NT, MG>(Mop, S.nv, S.part);
            __syncthreads();
            // P4: warp 0 combines and decides the step
            if (warp == 0) {
                double d2 = 0.0, nn = 0.0, sp = 0.0, t1 = INFINITY;
                int l = -1;
                for (int j = lane; j < nV; j += 32) {
                    const double yj = S.part[j] + S.part[nV + j] + S.part[2 * nV + j];
                    S.y[j] = yj;
                    const double nj = S.nv[j];
                    nn += nj * nj;
                    sp += nj * S.x[j];
                    if (j >= q) d2 += yj * yj;
                    else if (yj > 1e-13) {
                        const double tj = S.lam[j] / yj;
                        if (tj < t1) { t1 = tj; l = j; }
                    }
                }
                d2 = warp_sum(d2);
                nn = warp_sum(nn);
                sp = warp_sum(sp);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ot = __shfl_xor_sync(0xffffffffu, t1, o);
                    const int ol = __shfl_xor_sync(0xffffffffu, l, o);
                    if (ot < t1 || (ot == t1 && ol >= 0 && (l < 0 || ol < l))) { t1 = ot; l = ol; }
                }
                if (lane == 0) {
                    // rhs b of n'x >= b
                    double bnd;
                    if (pslot < nV) {
                        const double lb = (pslot < nU) ? P.u_lb[pslot % NU] : 0.0;
                        const double ub = (pslot < nU) ? P.u_ub[pslot % NU] : INFINITY;
                        bnd = pside < 0 ? lb : -ub;
                    } else {
                        bnd = pside < 0 ? S.rlo[pslot - nV] : -S.rup[pslot - nV];
                    }
                    sp -= bnd;
                    const bool lin_dep = !(d2 > 1e-13 * fmax(1.0, nn));
                    const double t2 = lin_dep ? INFINITY : (sp < 0.0 ? -sp / d2 : 0.0);
                    int type;
                    double t;
                    if (isinf(t1) && isinf(t2)) { type = STEP_INFEAS; t = 0.0; }
                    else if (t2 <= t1) { type = STEP_FULL; t = t2; }
                    else if (isinf(t2)) { type = STEP_DUAL; t = t1; }
                    else { type = STEP_PARTIAL; t = t1; }
                    const double delta = sqrt(d2);
                    const double yq = (q < nV) ? S.y[q] : 0.0;
                    const double sgd = (yq >= 0.0) ? delta : -delta;
                    S.sc[0] = t;
                    S.sc[1] = d2;
                    S.sc[2] = sgd;
                    S.sc[3] = 1.0 / (d2 + fabs(yq) * delta);   // beta = 2 / v'v
                    S.isc[3] = type;
                    S.isc[4] = l;
                    S.isc[5] += 1;
                    if (S.isc[5] > max_iter) { S.isc[3] = STEP_INFEAS; S.isc[6] = FSAE_EXIT_MAXITER; }
                    else if (type == STEP_INFEAS) S.isc[6] = FSAE_EXIT_INFEASIBLE;
                }
            }
            __syncthreads();
            const int type = S.isc[3];
            if (type == STEP_INFEAS) break;
            const double t = S.sc[0];
            // P5: primal step direction z = J2 y2
            if (type != STEP_DUAL) {
                matvec_N_parts<D, NT, MG>(Mop, S.y, q, S.part);
                __syncthreads();
                const double d2 = S.sc[1], sgd = S.sc[2];
                for (int i = tid; i < nV; i += NT) {
                    const double zi = S.part[i] + S.part[nV + i] + S.part[2 * nV + i];
                    S.x[i] += t * zi;
                    if (type == STEP_FULL) {
                        S.kv[i] = zi / d2;
                        S.wv[i] = zi + sgd * ldM<MG>(&Mop[i * LD + q]);
                    }
                }
            }
            for (int j = tid; j < q; j += NT) S.lam[j] -= t * S.y[j];
            if (tid == 0) S.sc[5] += t;
            __syncthreads();
            if (type == STEP_FULL) {
                // P6a: add p.  K1 <- K1 - k r',  J2 <- J2 (I - beta v v'), column q <- k
                const double beta = S.sc[3], sgd = S.sc[2];
                for (int tt = tid; tt < nV * nV; tt += NT) {
                    const int i = tt / nV, j = tt - i * nV;
                    double m = ldM<MG>(&Mop[i * LD + j]);
                    if (j < q) m -= S.kv[i] * S.y[j];
                    else if (j == q) m = S.kv[i];
                    else m -= beta * S.wv[i] * S.y[j];
                    Mop[i * LD + j] = m;
                }
                (void)sgd;
                if (tid == 0) {
                    S.act[q] = pslot * 2 + (pside > 0 ? 1 : 0);
                    S.lam[q] = S.sc[5];
                    S.status[pslot] = (int8_t)pside;
                    S.isc[0] = q + 1;
                }
                ++n_add;
                done_p = true;
                __syncthreads();
            } else {
                // P6b: drop active constraint l
                drop_column(S.isc[4]);
            }
        }
        if (S.isc[3] == STEP_INFEAS) break;
    }
    __syncthreads();

    // ---------------------------------------------------------------- outputs
    const int q = S.isc[0];
    // fval = 1/2 x'Hx + g'x + const  (ltvmpc_*_curvilinear.m:60), H without the flat_eps entries
    symv_packed<D, NT>(S.Hp, S.x, S.wv);
    __syncthreads();
    if (warp == 0) {
        double acc = 0.0;
        for (int i = lane; i < nV; i += 32) {
            double hx = S.wv[i];
            if (i >= nU) hx -= P.flat_eps * S.x[i];
            acc += S.x[i] * (0.5 * hx + S.g[i]);
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            a.fval[b] = acc + S.sc[7];
            a.exitflag[b] = S.isc[6];
            if (a.iters) a.iters[b] = S.isc[5];
            if (a.counters) {
                atomicAdd(a.counters + 0, n_add);
                atomicAdd(a.counters + 1, n_drop);
                atomicAdd(a.counters + 2, (unsigned long long)S.isc[7]);
            }
        }
    }
    for (int j = tid; j < nU; j += NT) a.u_opt[(size_t)b * nU + j] = S.x[j];
    for (int j = tid; j < NS; j += NT) a.slack_opt[(size_t)b * NS + j] = S.x[nU + j];
    // x_opt = A_bar x0 + B_bar u + d_bar = xf + (zero-state response to u)
    if (warp == 1 && lane == 0) {
        double xz[NX], xn[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) xz[i] = 0.0;
        double* gxo = a.x_opt + (size_t)b * NX * N;
        for (int k = 0; k < N; ++k) {
#pragma unroll
            for (int i = 0; i < NX; ++i) xn[i] = xz[i];
#pragma unroll
            for (int i = 0; i < C::NREAL; ++i) {
                double acc = 0.0;
#pragma unroll
                for (int c = 0; c < NX; ++c) acc += S.Ad[(k * C::NREAL + i) * NX + c] * xz[c];
                xn[C::real_state(i)] = acc;
            }
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                double acc = xn[i];
#pragma unroll
                for (int c = 0; c < NU; ++c) acc += S.B1[i * NU + c] * S.x[k * NU + c];
                xz[i] = acc;
                gxo[k * NX + i] = S.xf[k * NX + i] + acc;
            }
        }
    }
    if (a.wsB) {
        for (int j = tid; j < nV; j += NT) a.wsB[(size_t)b * nV + j] = S.status[j];
    }
    if (a.wsC) {
        int8_t* w = a.wsC + (size_t)b * C::n_ref_rows(N);
        for (int j = tid; j < C::n_ref_rows(N); j += NT) w[j] = 0;
        __syncthreads();
        for (int j = tid; j < q; j += NT) {
            const int code = S.act[j], slot = code >> 1, side = (code & 1) ? +1 : -1;
            if (slot >= nV) {
                const int rr = slot - nV, r = rr / N, k = rr - r * N;
                w[C::ref_row(r, k, side, N)] = (int8_t)side;
            }
        }
    }
}

}  // namespace fsae
