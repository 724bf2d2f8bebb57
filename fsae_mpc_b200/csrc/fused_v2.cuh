// Fused per-problem LTV-MPC step, version 2 (the product kernel).
//
// One CTA (8 warps) per problem.  The nV x nV dual active-set operator M = [K1 | J2] never
// touches shared memory: it lives in REGISTERS, tiled so that
//     warp w  owns rows    w*RPW .. w*RPW+RPW-1           (RPW = ceil(nV/8)  = 11 for nV = 81)
//     lane l  owns columns l, l+32, l+64                   (CS  = ceil(nV/32) = 3)
// i.e. every thread holds an RPW x CS tile (33 doubles).  With that layout
//   * y = M'n     : 33 FMAs per thread + ONE cross-warp sum through shared memory,
//   * z = J2 y2   : 33 FMAs per thread + an in-warp reduce-scatter (shuffles only),
//   * rank-1 update of M : 33 FMAs per thread, no communication,
// and every warp derives step lengths / add-or-drop decisions redundantly from the same
// data, so an iteration needs 4 block barriers and no serial "warp 0 decides" section.
// The condensed Hessian is accumulated, factorised (LDL' by symmetric elimination) and
// inverted in the same register tiles, so the factor goes H -> J = L^-T without leaving
// the register file.
//
// Pipeline and reference mapping: see fused_v1.cuh header (same stages).
#pragma once
#include "cons.cuh"
#include "fused_v1.cuh"   // BatchArgs, Dims, STEP_* enums, warp_sum

namespace fsae {

template <class Model, int N, int NW_ = 8>
struct CfgV2 {
    using D = Dims<Model, N>;
    static constexpr int NW = NW_, NT = 32 * NW_;
    static constexpr int RPW = (D::nV + NW - 1) / NW;     // rows per warp
    static constexpr int RP = RPW * NW;                     // padded rows
    static constexpr int CS = (D::nV + 31) / 32;            // column slots per lane
    static constexpr int CP = CS * 32;                      // padded columns
    static constexpr int RH = (RPW <= 8) ? 8 : 16;          // reduce-scatter width (pow2 >= RPW)
    static_assert(RPW <= 16, "reduce-scatter network supports up to 16 rows per warp");
    static_assert(D::nV < CP, "need one spare padded column for the piggy-backed scalar");
};

template <class Model, int N, int NW_ = 8>
struct SmemV2 {
    using D = Dims<Model, N>;
    using C = Cons<Model>;
    using G = CfgV2<Model, N, NW_>;
    double Bf[C::NREAL * D::NPK];      // packed B_bar rows of the "real" states (kept to the end)
    double Hp[D::HP];                  // packed lower triangle of H (drops, refresh, fval)
    double Ad[N * C::NREAL * D::NX];
    double B1[D::NX * D::NU];
    double xf[N * D::NX];
    double xl[N * D::NX];
    double ul[N * D::NU];
    double pc[N * C::NPC];
    double g0[N * C::NG0];
    double cg[C::NCG];
    double rlo[D::NROWS], rup[D::NROWS];
    double rn2[D::NROWS];              // squared norm of each row's normal (linear-dependence test)
    alignas(16) double x[G::RP];
    double g[G::RP];
    double ypart[2][G::NW][G::CP];     // cross-warp partial sums of M'v (double-buffered)
    double colk[2][G::RP];             // column broadcast (factorisation / column q / drop column)
    double rowv[G::RP];                // per-warp row vector scratch (gradient / H k)
    double nvec[G::RP];                // per-warp rows of the normal of the constraint being added
    double zrow[G::RP];                // per-warp reduced z
    double wpart[3][G::RP];            // symv partials
    double dvec[G::RP];                // LDL' pivots
    double dd[N * D::NX];
    double red_val[2][G::NW];
    double scal[8];                    // 0 cost const
    int red_idx[2][G::NW];
    int act[D::nV];
    int8_t status[D::NSLOT + 8];
};

// ---- in-warp reduce-scatter of RH values: afterwards lane l (and l^1) hold the warp-wide
// sum of entry (l >> 1) [RH = 16] or (l >> 2) [RH = 8].
template <int RH>
__device__ __forceinline__ double warp_reduce_scatter(double (&v)[RH]) {
    const int lane = threadIdx.x & 31;
    if constexpr (RH == 16) {
        {
            const bool hi = lane & 16;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const double send = hi ? v[i] : v[i + 8];
                const double keep = hi ? v[i + 8] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
        }
        {
            const bool hi = lane & 8;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double send = hi ? v[i] : v[i + 4];
                const double keep = hi ? v[i + 4] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
        }
        {
            const bool hi = lane & 4;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const double send = hi ? v[i] : v[i + 2];
                const double keep = hi ? v[i + 2] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
        }
        {
            const bool hi = lane & 2;
            const double send = hi ? v[0] : v[1];
            const double keep = hi ? v[1] : v[0];
            v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
        return v[0];
    } else {
        {
            const bool hi = lane & 16;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double send = hi ? v[i] : v[i + 4];
                const double keep = hi ? v[i + 4] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
        }
        {
            const bool hi = lane & 8;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const double send = hi ? v[i] : v[i + 2];
                const double keep = hi ? v[i + 2] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
        }
        {
            const bool hi = lane & 4;
            const double send = hi ? v[0] : v[1];
            const double keep = hi ? v[1] : v[0];
            v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
        return v[0];
    }
}
template <int RH>
__device__ __forceinline__ int rs_row_of_lane() {
    const int lane = threadIdx.x & 31;
    return RH == 16 ? (lane >> 1) : (lane >> 2);
}
template <int RH>
__device__ __forceinline__ bool rs_is_writer() {
    const int lane = threadIdx.x & 31;
    return RH == 16 ? ((lane & 1) == 0) : ((lane & 3) == 0);
}

template <class Model, int N, int MINB, int NW_ = 8>
__global__ void __launch_bounds__(32 * NW_, MINB) ltvmpc_fused_v2_kernel(BatchArgs a) {
    using D = Dims<Model, N>;
    using C = Cons<Model>;
    using G = CfgV2<Model, N, NW_>;
    using S_t = SmemV2<Model, N, NW_>;
    constexpr int NX = D::NX, NU = D::NU, NS = D::NS, nU = D::nU, nV = D::nV;
    constexpr int NT = G::NT, NW = G::NW, RPW = G::RPW, CS = G::CS, CP = G::CP, RP = G::RP, RH = G::RH;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    S_t& S = *reinterpret_cast<S_t*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    if (b >= a.B) return;
    const fsae_params& P = a.params[a.param_id ? a.param_id[b] : 0];
    const DevTrack tr = a.tracks[a.track_id ? a.track_id[b] : 0];
    const double dt = a.dt;
    const int row0 = warp * RPW;            // first row of this warp

    // ---------------------------------------------------------------- load
    {
        const double* gxl = a.x_lin + (size_t)b * NX * N;
        const double* gul = a.u_lin + (size_t)b * NU * N;
        for (int i = tid; i < NX * N; i += NT) S.xl[i] = gxl[i];
        for (int i = tid; i < NU * N; i += NT) S.ul[i] = gul[i];
        for (int i = tid; i < D::NSLOT; i += NT) S.status[i] = 0;
        for (int i = tid; i < RP; i += NT) { S.x[i] = 0.0; S.g[i] = 0.0; S.rowv[i] = 0.0; S.nvec[i] = 0.0; S.zrow[i] = 0.0; S.colk[0][i] = 0.0; S.colk[1][i] = 0.0; }
    }
    __syncthreads();

    // ---------------------------------------------------------------- linearise + discretise
    if (tid < N) {
        const int k = tid;
        double Ac[NX * NX], Bc_[NX * NU], dc[NX];
        linearise_step<Model>(P.lin_scheme, S.xl + k * NX, S.ul + k * NU, dt, tr, P, Ac, Bc_, dc);
#pragma unroll
        for (int i = 0; i < C::NREAL; ++i) {
            const int r = C::real_state(i);
#pragma unroll
            for (int c = 0; c < NX; ++c)
                S.Ad[(k * C::NREAL + i) * NX + c] = Ac[r * NX + c] * dt + (r == c ? 1.0 : 0.0);
        }
#pragma unroll
        for (int r = 0; r < NX; ++r) S.dd[k * NX + r] = dc[r] * dt;
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
    if (warp == NW - 1) {
        if (lane == 0) {
            double xp[NX], xn[NX];
            const double* gx0 = a.x0 + (size_t)b * NX;
#pragma unroll
            for (int i = 0; i < NX; ++i) xp[i] = gx0[i];
            for (int k = 0; k < N; ++k) {
#pragma unroll
                for (int i = 0; i < NX; ++i) xn[i] = xp[i];
#pragma unroll
                for (int i = 0; i < C::NREAL; ++i) {
                    double acc = 0.0;
#pragma unroll
                    for (int c = 0; c < NX; ++c) acc += S.Ad[(k * C::NREAL + i) * NX + c] * xp[c];
                    xn[C::real_state(i)] = acc;
                }
#pragma unroll
                for (int i = 0; i < NX; ++i) {
                    xn[i] += S.dd[k * NX + i];
                    S.xf[k * NX + i] = xn[i];
                    xp[i] = xn[i];
                }
            }
        }
    } else {
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
                for (int ii = 0; ii < C::NREAL; ++ii) S.Bf[ii * D::NPK + D::pk(k, t)] = v[C::real_state(ii)];
            }
        }
    }
    __syncthreads();

    // ---------------------------------------------------------------- g, bounds, row norms, cost const
    {
        double* e = S.dd;      // tracking error overwrites dd (dead after the free response)
        const double* gxr = a.x_ref + (size_t)b * NX * N;
        for (int i = tid; i < NX * N; i += NT) e[i] = S.xf[i] - gxr[i];
        __syncthreads();
        for (int j = tid; j < nV; j += NT) {
            double acc = 0.0;
            if (j < nU) {
                const int sj = j / NU, cj = j - sj * NU;
                for (int k = sj; k < N; ++k) {
#pragma unroll
                    for (int ii = 0; ii < C::NREAL; ++ii) {
                        const int r = C::real_state(ii);
                        const double q = (k == N - 1) ? P.Q_terminal[r] : P.Q[r];
                        acc += q * S.Bf[ii * D::NPK + D::pk(k, j)] * e[k * NX + r];
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
        if (warp == 0) {
            double acc = 0.0;
            for (int i = lane; i < NX * N; i += 32) {
                const int k = i / NX, r = i - k * NX;
                const double q = (k == N - 1) ? P.Q_terminal[r] : P.Q[r];
                acc += q * e[i] * e[i];
            }
            acc = warp_sum(acc);
            if (lane == 0) S.scal[0] = acc;
        }
        for (int t = tid; t < D::NROWS; t += NT) {
            const int r = t / N, k = t - r * N;
            double lo, up;
            C::row_bounds(r, S.xf + k * NX, S.xl + k * NX, S.ul + k * NU, S.pc + k * C::NPC, S.g0 + k * C::NG0, S.cg, P, lo, up);
            S.rlo[t] = lo;
            S.rup[t] = up;
            // squared norm of the row normal (u part + slack entry); only a scale for the
            // linear-dependence threshold, so large row sets use 1
            const double* pc = S.pc + k * C::NPC;
            double n2 = 1.0;
            if (D::NROWS <= 256) {
                n2 = (C::row_slack(r) >= 0) ? 1.0 : 0.0;
                for (int j = 0; j < NU * (k + 1); ++j) {
                    const int step = j / NU, uc = j - step * NU;
                    double v = 0.0;
#pragma unroll
                    for (int c = 0; c < C::NCR; ++c) v += C::row_coef(r, c, pc, S.cg) * S.Bf[C::cons_real(c) * D::NPK + D::pk(k, j)];
#pragma unroll
                    for (int c = 0; c < C::NINT; ++c)
                        if (C::int_ucol(c) == uc) v += C::row_coef(r, C::NCR + c, pc, S.cg) * dt;
                    if (step == k) v += C::row_ucoef(r, uc, pc, S.cg);
                    n2 += v * v;
                }
            }
            S.rn2[t] = n2;
        }
    }

    // ---------------------------------------------------------------- H in register tiles
    // generate_qp.m:29  H = 2 (B' Qbar B + Rbar) accumulated as a sum of rank-1 terms
    // q_{k,c} b_{k,c} b_{k,c}' over the rows (k, c) of B_bar, directly into the tile layout.
    double m[RPW][CS];
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
        for (int s = 0; s < CS; ++s) m[r][s] = 0.0;
    {
        // rows of this warp span controls row0 .. row0+RPW-1; a B_bar row (k, .) is nonzero on
        // controls j < NU (k+1): skip rows (k) that cannot touch this warp's tile rows.
        const int kmin = (row0 < nU) ? row0 / NU : N;
        for (int k = kmin; k < N; ++k) {
            const int len = NU * (k + 1);
#pragma unroll
            for (int ii = 0; ii < C::NREAL; ++ii) {
                const int rs = C::real_state(ii);
                const double qk = 2.0 * ((k == N - 1) ? P.Q_terminal[rs] : P.Q[rs]);
                if (qk == 0.0) continue;
                const double* brow = S.Bf + ii * D::NPK + D::pk(k, 0);
                double bj[CS];
#pragma unroll
                for (int s = 0; s < CS; ++s) {
                    const int j = lane + 32 * s;
                    bj[s] = (j < len) ? brow[j] * qk : 0.0;
                }
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    const int i = row0 + r;
                    const double bi = (i < len) ? brow[i] : 0.0;
#pragma unroll
                    for (int s = 0; s < CS; ++s) m[r][s] += bi * bj[s];
                }
            }
        }
        // integrator states (exact prefix rows: dt on their control up to step k) and R
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int i = row0 + r;
#pragma unroll
            for (int s = 0; s < CS; ++s) {
                const int j = lane + 32 * s;
                if (i < nU && j < nU) {
                    const int si = i / NU, ci = i - si * NU, sj = j / NU, cj = j - sj * NU;
                    const int sm = si > sj ? si : sj;
#pragma unroll
                    for (int ii = 0; ii < C::NINT; ++ii) {
                        if (C::int_ucol(ii) == ci && cj == ci) {
                            const int rs = C::int_state(ii);
                            m[r][s] += 2.0 * dt * dt * (P.Q[rs] * (double)(N - 1 - sm) + P.Q_terminal[rs]);
                        }
                    }
                    if (i == j) m[r][s] += 2.0 * P.R[ci];
                }
            }
        }
    }
    // packed copy for later symv's (slack diagonal gets flat_eps), optional debug tap
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int i = row0 + r;
#pragma unroll
        for (int s = 0; s < CS; ++s) {
            const int j = lane + 32 * s;
            if (i < nV && j <= i) S.Hp[D::hp(i, j)] = (i >= nU) ? (i == j ? P.flat_eps : 0.0) : m[r][s];
        }
    }
    if (a.dbg_H) {
        double* gH = a.dbg_H + (size_t)b * nV * nV;
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int i = row0 + r;
#pragma unroll
            for (int s = 0; s < CS; ++s) {
                const int j = lane + 32 * s;
                if (i < nV && j < nV) gH[(size_t)j * nV + i] = (i < nU && j < nU) ? m[r][s] : 0.0;
            }
        }
    }
    __syncthreads();
    if (a.dbg_g) {
        double* gg = a.dbg_g + (size_t)b * nV;
        for (int t = tid; t < nV; t += NT) gg[t] = S.g[t];
    }

    // ---------------------------------------------------------------- factor in registers
    // Symmetric elimination H_uu -> D with the column operations accumulated in place:
    // after step k, rows <= k of columns > k hold J_unit = (L1^-T) entries, the trailing
    // block holds the (full, symmetric) Schur complement.  One column broadcast + one
    // barrier per step; afterwards scale columns by d^-1/2:  J = L^-T, J J' = H_uu^-1.
    for (int k = 0; k < nU; ++k) {
        const int ks = k >> 5, kl = k & 31, buf = k & 1;
        if (lane == kl) {
#pragma unroll
            for (int s = 0; s < CS; ++s)
                if (s == ks) {
#pragma unroll
                    for (int r = 0; r < RPW; ++r) S.colk[buf][row0 + r] = m[r][s];
                }
        }
        __syncthreads();
        const double piv = S.colk[buf][k];
        const double rp = 1.0 / piv;
        if (tid == 0) S.dvec[k] = piv;
        double lj[CS];
#pragma unroll
        for (int s = 0; s < CS; ++s) {
            const int j = lane + 32 * s;
            lj[s] = (j > k && j < nU) ? S.colk[buf][j] * rp : 0.0;     // symmetric: W[k][j] = W[j][k]
        }
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const double vr = S.colk[buf][row0 + r];
#pragma unroll
            for (int s = 0; s < CS; ++s) m[r][s] = fma(-vr, lj[s], m[r][s]);
        }
        const int kr = k - row0;                    // warp-uniform: does this warp own the pivot row?
        if (kr >= 0 && kr < RPW) {
#pragma unroll
            for (int r = 0; r < RPW; ++r)
                if (r == kr) {
#pragma unroll
                    for (int s = 0; s < CS; ++s) {
                        const int j = lane + 32 * s;
                        if (j > k && j < nU) m[r][s] = -lj[s];
                    }
                }
        }
    }
    __syncthreads();
    double dcol[CS];                       // pivots of this lane's columns
#pragma unroll
    for (int s = 0; s < CS; ++s) {
        const int j = lane + 32 * s;
        dcol[s] = (j < nU) ? S.dvec[j] : 1.0;
    }
    // scale, clear the dead lower part, lay out M = [K1 (slack unit columns) | J2]:
    // column c of M:  c < NS -> e_{nU+c};  c >= NS -> J column (c - NS).  The tile holds J in
    // columns 0..nU-1, so shift columns right by NS through shared memory-free lane rotation:
    // instead of moving data we keep J where it is and put the slack columns LAST:
    // M columns 0..nU-1 = J2, columns nU..nV-1 = K1 slack columns.  The working set is then
    // "the last q_s columns + ..." -- to keep the [K1 | J2] convention (K1 first) we store the
    // column permutation implicitly: see `colperm` below.
#pragma unroll
    for (int s = 0; s < CS; ++s) {
        const int j = lane + 32 * s;
        const double sc = (j < nU) ? rsqrt(dcol[s]) : 0.0;
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int i = row0 + r;
            double v = 0.0;
            if (j < nU && i < nU) v = (i < j) ? m[r][s] * sc : (i == j ? sc : 0.0);
            m[r][s] = v;
        }
    }
    // Move the nU J-columns from positions 0..nU-1 to NS..nV-1 (K1 first): rotate columns right
    // by NS lanes.  Column j comes from column j-NS: lane (j-NS)&31, slot (j-NS)>>5.
    {
        double t[RPW][CS];
#pragma unroll
        for (int s = 0; s < CS; ++s) {
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                // value of column (lane + 32 s - NS): same slot if lane >= NS else previous slot
                const double same = __shfl_sync(0xffffffffu, m[r][s], (lane - NS) & 31);
                const double prev = (s > 0) ? __shfl_sync(0xffffffffu, m[r][s - 1], (lane - NS) & 31) : 0.0;
                t[r][s] = (lane >= NS) ? same : prev;
            }
        }
#pragma unroll
        for (int s = 0; s < CS; ++s) {
            const int j = lane + 32 * s;
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const int i = row0 + r;
                m[r][s] = (j < NS) ? ((i == nU + j) ? 1.0 : 0.0) : (j < nV ? t[r][s] : 0.0);
            }
        }
    }
    // working set: slack lower bounds active, multiplier = R_soft
    int q = NS;
    double lam[CS];
#pragma unroll
    for (int s = 0; s < CS; ++s) {
        const int j = lane + 32 * s;
        lam[s] = (j < NS) ? S.g[nU + j] : 0.0;
    }
    if (tid < NS) {
        S.act[tid] = (nU + tid) * 2;
        S.status[nU + tid] = -1;
    }

    int ybuf = 0;
    // y = M' v for a row vector v held per warp in S.rowv; returns y for this lane's columns
    // (identical in every warp).  The spare padded column CP-1 carries sum_i extra_i.
    auto matvec_T = [&](const double* rowvec, double extra, double (&y)[CS], double& extra_sum) {
        double yp[CS];
#pragma unroll
        for (int s = 0; s < CS; ++s) yp[s] = 0.0;
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const double v = rowvec[row0 + r];
#pragma unroll
            for (int s = 0; s < CS; ++s) yp[s] += m[r][s] * v;
        }
        if (lane == 31) yp[CS - 1] = extra;          // column CP-1 is padding (nV < CP)
#pragma unroll
        for (int s = 0; s < CS; ++s) S.ypart[ybuf][warp][lane + 32 * s] = yp[s];
        __syncthreads();
#pragma unroll
        for (int s = 0; s < CS; ++s) {
            double acc = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) acc += S.ypart[ybuf][w][lane + 32 * s];
            y[s] = acc;
        }
        ybuf ^= 1;
        extra_sum = __shfl_sync(0xffffffffu, y[CS - 1], 31);
        if (lane == 31) y[CS - 1] = 0.0;
    };
    // z = sum_{j >= q0} M[:, j] y_j for this warp's rows -> S.zrow (per warp), after __syncwarp
    auto matvec_N = [&](const double (&y)[CS], int q0) {
        double zp[RH];
#pragma unroll
        for (int r = 0; r < RH; ++r) zp[r] = 0.0;
#pragma unroll
        for (int s = 0; s < CS; ++s) {
            const int j = lane + 32 * s;
            const double yj = (j >= q0 && j < nV) ? y[s] : 0.0;
#pragma unroll
            for (int r = 0; r < RPW; ++r) zp[r] += m[r][s] * yj;
        }
        const double zr = warp_reduce_scatter<RH>(zp);
        const int rr = rs_row_of_lane<RH>();
        if (rs_is_writer<RH>() && rr < RPW) S.zrow[row0 + rr] = zr;
        __syncwarp();
    };
    // w = Hp * v (v in shared, full length) -> S.rowv (per-warp rows); includes a barrier
    auto symv_to_rowv = [&](const double* v, const double* addv) {
        constexpr int CH = (nV + 2) / 3;
        for (int t = tid; t < 3 * nV; t += NT) {
            const int pt = t / nV, i = t - pt * nV;
            const int j0 = pt * CH, j1 = (j0 + CH < nV) ? j0 + CH : nV;
            double acc = 0.0;
            for (int j = j0; j < j1; ++j) acc += ((j <= i) ? S.Hp[D::hp(i, j)] : S.Hp[D::hp(j, i)]) * v[j];
            S.wpart[pt][i] = acc;
        }
        __syncthreads();
        if (lane < RPW) {
            const int i = row0 + lane;
            if (i < nV) S.rowv[i] = S.wpart[0][i] + S.wpart[1][i] + S.wpart[2][i] + (addv ? addv[i] : 0.0);
        }
        __syncwarp();
    };

    __syncthreads();
    // ---------------------------------------------------------------- x0 = -J2 J2' g (slack at bound)
    {
        double y[CS], dummy;
        matvec_T(S.g, 0.0, y, dummy);
        matvec_N(y, q);
        if (lane < RPW) {
            const int i = row0 + lane;
            if (i < nU) S.x[i] = -S.zrow[i];
        }
    }
    __syncthreads();                           // x complete

    // ---------------------------------------------------------------- Goldfarb-Idnani, K-form
    const double tol = P.feas_tol;
    const int max_iter = P.max_iter;
    int iters = 0, exitflag = FSAE_EXIT_SOLVED, n_add = 0, n_drop = 0, n_refresh = 0;
    int rbuf = 0;
    while (true) {
        // P1: most violated inactive constraint side.  Threads [0, 4N): 4 lanes per horizon
        // step k compute the constraint-state perturbations xs[., k] = (B_bar_c x)[., k]
        // (packed rows, plus the exact prefix sums of the integrator states) and then split
        // that step's rows among themselves; threads [4N, 4N+nV): the variable bounds.
        // No shared-memory round trip and no barrier between evaluation and search.
        static_assert(4 * N + nV <= NT, "P1 thread map needs 4N + nV <= 256");
        static_assert(NU == 2, "paired (double2) row loads assume two controls per step");
        double best = 0.0;
        int best_i = 0x7fffffff;
        if (warp < (4 * N + 31) / 32) {
            const int k = tid >> 2, part = tid & 3;
            const bool valid = k < N;
            double acc[C::NXS];
#pragma unroll
            for (int c = 0; c < C::NXS; ++c) acc[c] = 0.0;
            if (valid) {
                const int len = NU * (k + 1);
                const int ch = (((len + 3) >> 2) + 1) & ~1;          // even chunk -> 16-byte aligned pairs
                const int j0 = part * ch, j1 = (j0 + ch < len) ? j0 + ch : len;
                for (int j = j0; j < j1; j += 2) {
                    const double2 xx = *reinterpret_cast<const double2*>(&S.x[j]);
#pragma unroll
                    for (int c = 0; c < C::NCR; ++c) {
                        const double2 bb = *reinterpret_cast<const double2*>(&S.Bf[C::cons_real(c) * D::NPK + D::pk(k, j)]);
                        acc[c] = fma(bb.x, xx.x, fma(bb.y, xx.y, acc[c]));
                    }
#pragma unroll
                    for (int ci = 0; ci < C::NINT; ++ci) acc[C::NCR + ci] += (C::int_ucol(ci) == 0) ? xx.x : xx.y;
                }
            }
#pragma unroll
            for (int c = 0; c < C::NXS; ++c) {
                acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 1);
                acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 2);
            }
            if (valid) {
#pragma unroll
                for (int ci = 0; ci < C::NINT; ++ci) acc[C::NCR + ci] *= dt;
                const double ua = S.x[NU * k];
                for (int r = part; r < C::NR; r += 4) {
                    const int rr = r * N + k, slot = nV + rr;
                    if (S.status[slot] != 0) continue;
                    const double rv = C::row_value(r, acc, S.pc + k * C::NPC, S.cg, ua);
                    const int sl = C::row_slack(r);
                    const double sv = sl >= 0 ? S.x[nU + sl] : 0.0;
                    const double vlo = rv + sv - S.rlo[rr];
                    const double vup = S.rup[rr] - rv + sv;
                    if (vlo < best) { best = vlo; best_i = slot * 2; }
                    if (vup < best) { best = vup; best_i = slot * 2 + 1; }
                }
            }
        } else {
            const int slot = tid - 4 * N;
            if (slot >= 0 && slot < nV && S.status[slot] == 0) {
                const double xv = S.x[slot];
                const double lb = (slot < nU) ? P.u_lb[slot % NU] : 0.0;
                const double ub = (slot < nU) ? P.u_ub[slot % NU] : INFINITY;
                const double vlo = xv - lb, vup = ub - xv;
                if (vlo < best) { best = vlo; best_i = slot * 2; }
                if (vup < best) { best = vup; best_i = slot * 2 + 1; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (ov < best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
        }
        if (lane == 0) { S.red_val[rbuf][warp] = best; S.red_idx[rbuf][warp] = best_i; }
        __syncthreads();
        double viol = S.red_val[rbuf][0];
        int pcode = S.red_idx[rbuf][0];
#pragma unroll
        for (int w = 1; w < NW; ++w) {
            const double ov = S.red_val[rbuf][w];
            const int oi = S.red_idx[rbuf][w];
            if (ov < viol || (ov == viol && oi < pcode)) { viol = ov; pcode = oi; }
        }
        rbuf ^= 1;

        if (!(viol < -tol)) {
            if (n_refresh >= 1) break;
            // refresh: Newton step on the active manifold + multipliers from stationarity
            ++n_refresh;
            symv_to_rowv(S.x, S.g);                         // rowv = H x + g
            double y[CS], dummy;
            matvec_T(S.rowv, 0.0, y, dummy);
            matvec_N(y, q);
            if (lane < RPW) {
                const int i = row0 + lane;
                if (i < nV) S.x[i] -= S.zrow[i];
            }
#pragma unroll
            for (int s = 0; s < CS; ++s) {
                const int j = lane + 32 * s;
                if (j < q) lam[s] = fmax(y[s], 0.0);
            }
            __syncthreads();                   // x complete
            continue;
        }
        const int pslot = pcode >> 1, pside = (pcode & 1) ? +1 : -1;
        double sp = viol;                                   // n'x - b  (< 0)
        double lam_p = 0.0;
        // P2: this warp's entries of the normal -> S.rowv
        if (lane < RPW) {
            const int i = row0 + lane;
            double v = 0.0;
            if (i < nV) {
                const double sg = pside < 0 ? 1.0 : -1.0;
                if (pslot < nV) {
                    v = (i == pslot) ? sg : 0.0;
                } else {
                    const int rr = pslot - nV, r = rr / N, k = rr - r * N;
                    if (i >= nU) {
                        const int sl = C::row_slack(r);
                        v = (sl >= 0 && i == nU + sl) ? 1.0 : 0.0;
                    } else {
                        const int step = i / NU, uc = i - step * NU;
                        if (step <= k) {
                            const double* pc = S.pc + k * C::NPC;
                            double acc = 0.0;
#pragma unroll
                            for (int c = 0; c < C::NCR; ++c)
                                acc += C::row_coef(r, c, pc, S.cg) * S.Bf[C::cons_real(c) * D::NPK + D::pk(k, i)];
#pragma unroll
                            for (int c = 0; c < C::NINT; ++c)
                                if (C::int_ucol(c) == uc) acc += C::row_coef(r, C::NCR + c, pc, S.cg) * dt;
                            if (step == k) acc += C::row_ucoef(r, uc, pc, S.cg);
                            v = sg * acc;
                        }
                    }
                }
            }
            S.nvec[row0 + lane] = v;
        }
        __syncwarp();
        const double nn = (pslot < nV) ? 1.0 : S.rn2[pslot - nV];

        bool failed = false;
        while (true) {
            if (++iters > max_iter) { exitflag = FSAE_EXIT_MAXITER; failed = true; break; }
            // P3: y = M' n
            double y[CS], dummy;
            matvec_T(S.nvec, 0.0, y, dummy);
            // P4 (every warp, redundantly): step lengths
            double d2 = 0.0, t1 = INFINITY;
            int l = -1;
#pragma unroll
            for (int s = 0; s < CS; ++s) {
                const int j = lane + 32 * s;
                if (j >= q && j < nV) d2 += y[s] * y[s];
                else if (j < q && y[s] > 1e-13) {
                    const double tj = lam[s] / y[s];
                    if (tj < t1) { t1 = tj; l = j; }
                }
            }
            d2 = warp_sum(d2);
            {
                double tm = t1;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) tm = fmin(tm, __shfl_xor_sync(0xffffffffu, tm, o));
                const unsigned who = __ballot_sync(0xffffffffu, l >= 0 && t1 == tm);
                if (who) l = __shfl_sync(0xffffffffu, l, __ffs(who) - 1);
                t1 = tm;
            }
            const bool lin_dep = !(d2 > 1e-13 * fmax(1.0, nn));
            const double t2 = lin_dep ? INFINITY : (sp < 0.0 ? -sp / d2 : 0.0);
            if (isinf(t1) && isinf(t2)) { exitflag = FSAE_EXIT_INFEASIBLE; failed = true; break; }
            const bool full = (t2 <= t1);
            const bool primal = !isinf(t2);
            const double t = full ? t2 : t1;
            // P5: z = J2 y2, x += t z
            if (primal) {
                matvec_N(y, q);
                if (lane < RPW) {
                    const int i = row0 + lane;
                    if (i < nV) S.x[i] += t * S.zrow[i];
                }
                sp += t * d2;
            }
            if (full) {
                // bookkeeping of the add happens BEFORE the barrier that publishes x, so the
                // next search (which follows the register-only update below without another
                // barrier) sees a consistent x / status
                if (tid == 0) {
                    S.act[q] = pslot * 2 + (pside > 0 ? 1 : 0);
                    S.status[pslot] = (int8_t)pside;
                }
                __syncthreads();               // x complete
            }
#pragma unroll
            for (int s = 0; s < CS; ++s) {
                const int j = lane + 32 * s;
                if (j < q) lam[s] -= t * y[s];
            }
            lam_p += t;
            if (full) {
                // P6a: add p.  K1 <- K1 - k r', J2 <- J2 (I - beta v v'), column q <- k = z/d2
                const int qs = q >> 5, ql = q & 31;
                if (lane == ql) {
#pragma unroll
                    for (int s = 0; s < CS; ++s)
                        if (s == qs) {
#pragma unroll
                            for (int r = 0; r < RPW; ++r) S.colk[0][row0 + r] = m[r][s];
                        }
                }
                __syncwarp();
                const double delta = sqrt(d2);
                const double yq = __shfl_sync(0xffffffffu, (qs == 0 ? y[0] : (qs == 1 ? y[CS > 1 ? 1 : 0] : y[CS - 1])), ql);
                const double sgd = (yq >= 0.0) ? delta : -delta;
                const double beta = 1.0 / (d2 + fabs(yq) * delta);
                const double inv_d2 = 1.0 / d2;
                // new = c*cur - kr*ya - wr*yb with (c, ya, yb) = (1, y, 0) for j < q,
                // (0, -1, 0) for j == q, (1, 0, y) for j > q: no per-element selects
                double cc[CS], ya[CS], yb[CS];
#pragma unroll
                for (int s = 0; s < CS; ++s) {
                    const int j = lane + 32 * s;
                    cc[s] = (j == q) ? 0.0 : 1.0;
                    ya[s] = (j < q) ? y[s] : (j == q ? -1.0 : 0.0);
                    yb[s] = (j > q) ? y[s] : 0.0;
                }
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    const double zr = S.zrow[row0 + r];
                    const double kr = zr * inv_d2;
                    const double wr = (zr + sgd * S.colk[0][row0 + r]) * beta;
#pragma unroll
                    for (int s = 0; s < CS; ++s) m[r][s] = fma(-wr, yb[s], fma(-kr, ya[s], cc[s] * m[r][s]));
                }
#pragma unroll
                for (int s = 0; s < CS; ++s) {
                    const int j = lane + 32 * s;
                    if (j == q) lam[s] = lam_p;
                }
                ++q;
                ++n_add;
                __syncwarp();
                break;
            }
            // P6b: drop active constraint l (column l of K1)
            {
                const int ls = l >> 5, ll = l & 31;
                if (lane == ll) {
#pragma unroll
                    for (int s = 0; s < CS; ++s)
                        if (s == ls) {
#pragma unroll
                            for (int r = 0; r < RPW; ++r) S.colk[1][row0 + r] = m[r][s];
                        }
                }
                __syncthreads();                                 // k = M[:, l] visible block-wide
                symv_to_rowv(S.colk[1], nullptr);                // rowv = H k   (barrier inside)
                double kw = 0.0;
                if (lane < RPW && row0 + lane < nV) kw = S.colk[1][row0 + lane] * S.rowv[row0 + lane];
                kw = warp_sum(kw);
                double rp[CS], kHk;
                matvec_T(S.rowv, kw, rp, kHk);                   // rp_j = M[:,j]' (H k), kHk piggy-backed
                const double ik = 1.0 / kHk;
                const double rs = rsqrt(kHk);
                const int q1 = q - 1, q1s = q1 >> 5, q1l = q1 & 31;
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    const double kr = S.colk[1][row0 + r];
                    // K1 <- K1 + k r'^T with r' = -K1' H k / kHk  (columns j < q, j != l)
#pragma unroll
                    for (int s = 0; s < CS; ++s) {
                        const int j = lane + 32 * s;
                        if (j < q && j != l) m[r][s] -= kr * (rp[s] * ik);
                    }
                    // column q-1 -> column l ; column q-1 <- k / sqrt(kHk)
                    double last = 0.0;
#pragma unroll
                    for (int s = 0; s < CS; ++s)
                        if (s == q1s) last = m[r][s];
                    last = __shfl_sync(0xffffffffu, last, q1l);
#pragma unroll
                    for (int s = 0; s < CS; ++s) {
                        const int j = lane + 32 * s;
                        if (j == l && l != q1) m[r][s] = last;
                    }
#pragma unroll
                    for (int s = 0; s < CS; ++s) {
                        const int j = lane + 32 * s;
                        if (j == q1) m[r][s] = kr * rs;
                    }
                }
                double lam_last = 0.0;
#pragma unroll
                for (int s = 0; s < CS; ++s)
                    if (s == q1s) lam_last = lam[s];
                lam_last = __shfl_sync(0xffffffffu, lam_last, q1l);
#pragma unroll
                for (int s = 0; s < CS; ++s) {
                    const int j = lane + 32 * s;
                    if (j == l && l != q1) lam[s] = lam_last;
                    if (j == q1) lam[s] = 0.0;
                }
                if (tid == 0) {
                    S.status[S.act[l] >> 1] = 0;
                    S.act[l] = S.act[q1];
                }
                --q;
                ++n_drop;
            }
        }
        if (failed) break;
    }
    __syncthreads();

    // ---------------------------------------------------------------- outputs
    // fval = 1/2 x'Hx + g'x + const (ltvmpc_*_curvilinear.m:60); H without the flat_eps entries
    symv_to_rowv(S.x, nullptr);
    {
        double acc = 0.0;
        if (lane < RPW) {
            const int i = row0 + lane;
            if (i < nV) {
                double hx = S.rowv[i];
                if (i >= nU) hx -= P.flat_eps * S.x[i];
                acc = S.x[i] * (0.5 * hx + S.g[i]);
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) S.red_val[0][warp] = acc;
    }
    __syncthreads();
    if (tid == 0) {
        double f = S.scal[0];
#pragma unroll
        for (int w = 0; w < NW; ++w) f += S.red_val[0][w];
        a.fval[b] = f;
        a.exitflag[b] = exitflag;
        if (a.iters) a.iters[b] = iters;
        if (a.counters) {
            atomicAdd(a.counters + 0, (unsigned long long)n_add);
            atomicAdd(a.counters + 1, (unsigned long long)n_drop);
            atomicAdd(a.counters + 2, (unsigned long long)n_refresh);
        }
    }
    for (int j = tid; j < nU; j += NT) a.u_opt[(size_t)b * nU + j] = S.x[j];
    for (int j = tid; j < NS; j += NT) a.slack_opt[(size_t)b * NS + j] = S.x[nU + j];
    // x_opt = A_bar x0 + B_bar u + d_bar = xf + B_bar u  (ltvmpc_*_curvilinear.m:58)
    {
        double* gxo = a.x_opt + (size_t)b * NX * N;
        constexpr int NRR = C::NREAL * N;
        for (int base = 0; base < NRR; base += NT / 4) {
            const int rid = base + (tid >> 2), part = tid & 3;
            double acc = 0.0;
            int c = 0, k = 0;
            if (rid < NRR) {
                c = rid / N; k = rid - c * N;
                const double* row = S.Bf + c * D::NPK + D::pk(k, 0);
                const int len = NU * (k + 1);
                const int ch = (len + 3) >> 2;
                const int j0 = part * ch, j1 = (j0 + ch < len) ? j0 + ch : len;
                for (int j = j0; j < j1; ++j) acc += row[j] * S.x[j];
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            if (rid < NRR && part == 0) gxo[k * NX + C::real_state(c)] = S.xf[k * NX + C::real_state(c)] + acc;
        }
        for (int t = tid; t < C::NINT * N; t += NT) {
            const int ci = t / N, k = t - ci * N;
            const int uc = C::int_ucol(ci), rs = C::int_state(ci);
            double acc = 0.0;
            for (int i = 0; i <= k; ++i) acc += S.x[NU * i + uc];
            gxo[k * NX + rs] = S.xf[k * NX + rs] + acc * dt;
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
