// Register-tiled dual active-set core shared by the fused LTV-MPC kernels and the dense
// qpOASES drop-in kernel.
//
//   min 1/2 x'Hx + g'x   s.t. "slots" (variable bounds and general rows), each two-sided
//
// Goldfarb-Idnani (Math. Prog. 27, 1983) in operator form: the nV x nV matrix M = [K1 | J2]
//   J2 : H-orthonormal basis of the null space of the working set      (J2' H J2 = I, N' J2 = 0)
//   K1 : multiplier operator H^-1 N (N' H^-1 N)^-1                       (N' K1 = I, J2' H K1 = 0)
// lives in REGISTERS: warp w owns rows w*RPW .. w*RPW+RPW-1, lane l owns columns l, l+32, ...
// (an RPW x CS tile per thread).  One iteration:
//   y = M'n (one cross-warp sum through smem), step lengths and the add/drop decision computed
//   redundantly by every warp, z = J2 y2 (in-warp reduce-scatter), then either
//   add : Householder on J2 + rank-1 on K1 (3 FP64 ops per element, no communication), or
//   drop: K1 += k r'^T with r' = -K1' H k / k'Hk, column swap by shuffle.
// The problem-specific parts (how slots are evaluated and what their normals are) come from a
// policy object:  search(best, best_i),  normal_entry(slot, side, i),  norm2(slot).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace fsae {

// Phase timers for kernel tuning (scripts/phase_profile.py builds a separate .so with
// -DFSAE_PROFILE; the product build compiles them out).
#ifdef FSAE_PROFILE
__device__ unsigned long long g_phase_cycles[16];
#define PHASE_DECL long long ph_t_ = clock64(); unsigned long long ph_acc_[12] = {0}
#define PHASE(i) do { const long long n_ = clock64(); ph_acc_[i] += (unsigned long long)(n_ - ph_t_); ph_t_ = n_; } while (0)
#define PHASE_FLUSH do { if (threadIdx.x == 0) for (int i_ = 0; i_ < 12; ++i_) atomicAdd(&g_phase_cycles[i_], ph_acc_[i_]); } while (0)
#else
#define PHASE_DECL
#define PHASE(i)
#define PHASE_FLUSH
#endif

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- in-warp reduce-scatter of RH values: afterwards lane l (and its RH-group partners) hold
// the warp-wide sum of entry (l >> 1) [RH = 16] or (l >> 2) [RH = 8].
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

// Tile geometry for at most NVMAX variables and NW warps per CTA.
template <int NVMAX, int NW_ = 8>
struct GiCfg {
    static constexpr int NW = NW_, NT = 32 * NW_;
    static constexpr int RPW = (NVMAX + NW - 1) / NW;      // rows per warp
    static constexpr int RP = RPW * NW;                      // padded rows
    static constexpr int CS = (NVMAX + 32) / 32;             // column slots per lane (one spare column)
    static constexpr int CP = CS * 32;                       // padded columns
    static constexpr int RH = (RPW <= 8) ? 8 : 16;           // reduce-scatter width
    static constexpr int HP = NVMAX * (NVMAX + 1) / 2;       // packed lower triangle
    static_assert(RPW <= 16, "reduce-scatter network supports up to 16 rows per warp");
    static_assert(NVMAX < CP, "need one spare padded column for the piggy-backed scalar");
    __host__ __device__ static constexpr int hp(int i, int j) { return i * (i + 1) / 2 + j; }   // i >= j
};

// Shared-memory working set of the core.
template <class G, int NSLOT>
struct GiSm {
    alignas(16) double x[G::RP];
    double g[G::RP];
    double Hp[G::HP];                  // packed lower triangle of H (drops, refresh, fval)
    double ypart[2][G::NW][G::CP];     // cross-warp partial sums of M'v (double-buffered)
    double colk[2][G::RP];             // column broadcast (factorisation / column q / drop column)
    double rowv[G::RP];                // per-warp row vector scratch (gradient / H k)
    double nvec[G::RP];                // per-warp rows of the normal of the constraint being added
    double zrow[G::RP];                // per-warp reduced z
    double wpart[3][G::RP];            // symv partials
    double dvec[G::RP];                // LDL' pivots
    double red_val[2][G::NW];
    int red_idx[2][G::NW];
    int act[G::RP];                    // slot*2 + (side > 0) of working-set column j
    int8_t status[NSLOT + 8];          // -1 / 0 / +1 per slot
};

struct GiStats {
    int iters, exitflag, n_add, n_drop, n_refresh;
};

enum { GI_EXIT_SOLVED = 0, GI_EXIT_MAXITER = 1, GI_EXIT_INFEASIBLE = -2 };

template <class G, class SM>
struct GiOps {
    static constexpr int NW = G::NW, NT = G::NT, RPW = G::RPW, CS = G::CS, RH = G::RH;

    // y = M' v for a row vector v (shared, per-warp rows); y for this lane's columns, identical
    // in every warp.  The spare padded column CP-1 carries sum over warps of `extra`.
    __device__ __forceinline__ static void matvec_T(SM& S, const double (&m)[RPW][CS], int& ybuf,
                                                    const double* rowvec, double extra,
                                                    double (&y)[CS], double& extra_sum) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, row0 = warp * RPW;
        double yp[CS];
#pragma unroll
        for (int s = 0; s < CS; ++s) yp[s] = 0.0;
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const double v = rowvec[row0 + r];
#pragma unroll
            for (int s = 0; s < CS; ++s) yp[s] += m[r][s] * v;
        }
        if (lane == 31) yp[CS - 1] = extra;
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
    }

    // z = sum_{q0 <= j < nV} M[:, j] y_j for this warp's rows -> S.zrow, visible after __syncwarp
    __device__ __forceinline__ static void matvec_N(SM& S, const double (&m)[RPW][CS], const double (&y)[CS],
                                                    int q0, int nV) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, row0 = warp * RPW;
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
        const int rr = (RH == 16) ? (lane >> 1) : (lane >> 2);
        const bool writer = (RH == 16) ? ((lane & 1) == 0) : ((lane & 3) == 0);
        if (writer && rr < RPW) S.zrow[row0 + rr] = zr;
        __syncwarp();
    }

    // rowv = Hp * v (+ addv) for this warp's rows; v full length in shared; one barrier inside
    __device__ __forceinline__ static void symv_to_rowv(SM& S, const double* v, const double* addv, int nV) {
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, row0 = warp * RPW;
        const int CH = (nV + 2) / 3;
        for (int t = tid; t < 3 * nV; t += NT) {
            const int pt = t / nV, i = t - pt * nV;
            const int j0 = pt * CH, j1 = (j0 + CH < nV) ? j0 + CH : nV;
            double acc = 0.0;
            for (int j = j0; j < j1; ++j) acc += ((j <= i) ? S.Hp[G::hp(i, j)] : S.Hp[G::hp(j, i)]) * v[j];
            S.wpart[pt][i] = acc;
        }
        __syncthreads();
        if (lane < RPW) {
            const int i = row0 + lane;
            if (i < nV) S.rowv[i] = S.wpart[0][i] + S.wpart[1][i] + S.wpart[2][i] + (addv ? addv[i] : 0.0);
        }
        __syncwarp();
    }

    // Symmetric elimination of the leading nC x nC block held in the tiles (the "curved"
    // variables) with the column operations accumulated in place, then the layout
    //   M = [ e_{nC}, .., e_{nC+ns-1} | J ],  J = L^-T (J J' = H_cc^-1), flat variables last.
    // The ns flat (zero-curvature) variables start with one bound in the working set:
    // q = ns, lam_j = |g_flat_j|.  One column broadcast + one barrier per elimination step.
    // Returns false (uniformly) if a pivot is not positive.
    __device__ static bool factor_and_layout(SM& S, double (&m)[RPW][CS], double (&lam)[CS], int& q,
                                             int nC, int ns, int nV) {
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, row0 = warp * RPW;
        bool ok = true;
        for (int k = 0; k < nC; ++k) {
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
            ok = ok && (piv > 0.0);
            const double rp = 1.0 / piv;
            if (tid == 0) S.dvec[k] = piv;
            double lj[CS];
#pragma unroll
            for (int s = 0; s < CS; ++s) {
                const int j = lane + 32 * s;
                lj[s] = (j > k && j < nC) ? S.colk[buf][j] * rp : 0.0;     // symmetric: W[k][j] = W[j][k]
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
                            if (j > k && j < nC) m[r][s] = -lj[s];
                        }
                    }
            }
        }
        __syncthreads();
        // scale columns by d^-1/2, clear the dead lower part
#pragma unroll
        for (int s = 0; s < CS; ++s) {
            const int j = lane + 32 * s;
            const double sc = (j < nC) ? rsqrt(S.dvec[j]) : 0.0;
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const int i = row0 + r;
                double v = 0.0;
                if (j < nC && i < nC) v = (i < j) ? m[r][s] * sc : (i == j ? sc : 0.0);
                m[r][s] = v;
            }
        }
        // rotate the J columns right by ns lanes (K1 first): column j comes from column j - ns
        {
            double t[RPW][CS];
#pragma unroll
            for (int s = 0; s < CS; ++s) {
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    const double same = __shfl_sync(0xffffffffu, m[r][s], (lane - ns) & 31);
                    const double prev = (s > 0) ? __shfl_sync(0xffffffffu, m[r][s - 1], (lane - ns) & 31) : 0.0;
                    t[r][s] = (lane >= ns) ? same : prev;
                }
            }
#pragma unroll
            for (int s = 0; s < CS; ++s) {
                const int j = lane + 32 * s;
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    const int i = row0 + r;
                    m[r][s] = (j < ns) ? ((i == nC + j) ? 1.0 : 0.0) : (j < nV ? t[r][s] : 0.0);
                }
            }
        }
        q = ns;
#pragma unroll
        for (int s = 0; s < CS; ++s) {
            const int j = lane + 32 * s;
            lam[s] = (j < ns) ? fabs(S.g[nC + j]) : 0.0;
        }
        return ok;
    }

    // x_c = -J2 J2' g for the curved variables (flat ones keep the bound the caller put in x)
    __device__ static void initial_point(SM& S, const double (&m)[RPW][CS], int& ybuf, int q, int nC, int nV) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, row0 = warp * RPW;
        double y[CS], dummy;
        matvec_T(S, m, ybuf, S.g, 0.0, y, dummy);
        matvec_N(S, m, y, q, nV);
        if (lane < RPW) {
            const int i = row0 + lane;
            if (i < nC) S.x[i] = -S.zrow[i];
        }
        __syncthreads();                           // x complete
    }

    // The dual active-set loop.  On entry: M, lam, q consistent with S.act/S.status, S.x the
    // minimiser on that working set, barrier passed.  On exit S.x is the solution (barrier passed).
    template <class Prob>
    __device__ static GiStats solve(const Prob& prob, SM& S, double (&m)[RPW][CS], double (&lam)[CS], int& q,
                                    int& ybuf, int nV, double tol, int max_iter) {
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, row0 = warp * RPW;
        GiStats st = {0, GI_EXIT_SOLVED, 0, 0, 0};
        int rbuf = 0;
        PHASE_DECL;
        while (true) {
            // P1: most violated inactive constraint side (policy evaluates its slots)
            double best = 0.0;
            int best_i = 0x7fffffff;
            PHASE(0);
            prob.search(best, best_i);
            PHASE(1);
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
            PHASE(2);

            if (!(viol < -tol)) {
                // the refresh repairs what chains of partial steps leave behind; a run of pure
                // full steps keeps x the exact working-set minimiser (to round-off)
                if (st.n_refresh >= 1 || st.n_drop == 0) break;
                // refresh: Newton step on the active manifold + multipliers from stationarity
                ++st.n_refresh;
                symv_to_rowv(S, S.x, S.g, nV);                       // rowv = H x + g
                double y[CS], dummy;
                matvec_T(S, m, ybuf, S.rowv, 0.0, y, dummy);
                matvec_N(S, m, y, q, nV);
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
            // P2: this warp's entries of the normal
            {
                const auto prep = prob.normal_prepare(pslot, pside);      // warp-uniform part
                if (lane < RPW) {
                    const int i = row0 + lane;
                    S.nvec[i] = (i < nV) ? prob.normal_entry(prep, i) : 0.0;
                }
            }
            __syncwarp();
            const double nn = prob.norm2(pslot);
            PHASE(3);

            bool failed = false;
            while (true) {
                if (++st.iters > max_iter) { st.exitflag = GI_EXIT_MAXITER; failed = true; break; }
                // P3: y = M' n.  For a variable bound n = +-e_p, so y is +-(row p of M): the owning
                // warp publishes its row, nobody multiplies or sums partials.
                double y[CS], dummy;
                if (prob.is_unit(pslot)) {
                    const int pr = pslot - row0;               // warp-uniform
                    if (pr >= 0 && pr < RPW) {
#pragma unroll
                        for (int r = 0; r < RPW; ++r)
                            if (r == pr) {
#pragma unroll
                                for (int s = 0; s < CS; ++s) S.ypart[ybuf][0][lane + 32 * s] = m[r][s];
                            }
                    }
                    __syncthreads();
                    const double sgn = pside < 0 ? 1.0 : -1.0;
#pragma unroll
                    for (int s = 0; s < CS; ++s) y[s] = sgn * S.ypart[ybuf][0][lane + 32 * s];
                    ybuf ^= 1;
                } else {
                    matvec_T(S, m, ybuf, S.nvec, 0.0, y, dummy);
                }
                PHASE(4);
                // P4 (every warp, redundantly): step lengths
                double d2 = 0.0, t1 = INFINITY;
                int l = -1;
#pragma unroll
                for (int s = 0; s < CS; ++s) {      // branch-free: every lane does the same work
                    const int j = lane + 32 * s;
                    const double yy = y[s];
                    const bool isJ = (j >= q) & (j < nV);
                    const bool cand = (j < q) & (yy > 1e-13);
                    d2 = fma(isJ ? yy : 0.0, yy, d2);
                    const double tj = cand ? lam[s] * __drcp_rn(cand ? yy : 1.0) : INFINITY;
                    const bool better = tj < t1;
                    t1 = better ? tj : t1;
                    l = better ? j : l;
                }
                d2 = warp_sum_d(d2);
                {
                    double tm = t1;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) tm = fmin(tm, __shfl_xor_sync(0xffffffffu, tm, o));
                    const unsigned who = __ballot_sync(0xffffffffu, l >= 0 && t1 == tm);
                    if (who) l = __shfl_sync(0xffffffffu, l, __ffs(who) - 1);
                    t1 = tm;
                }
                const bool lin_dep = !(d2 > 1e-13 * fmax(1.0, nn));
                const double inv_d2 = __drcp_rn(d2);
                const double t2 = lin_dep ? INFINITY : (sp < 0.0 ? -sp * inv_d2 : 0.0);
                if (isinf(t1) && isinf(t2)) { st.exitflag = GI_EXIT_INFEASIBLE; failed = true; break; }
                const bool full = (t2 <= t1);
                const bool primal = !isinf(t2);
                const double t = full ? t2 : t1;
                PHASE(5);
                // P5: z = J2 y2, x += t z
                if (primal) {
                    matvec_N(S, m, y, q, nV);
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
                PHASE(6);
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
                    const double delta = d2 * rsqrt(d2);
                    double yq_l = 0.0;
#pragma unroll
                    for (int s = 0; s < CS; ++s)
                        if (s == qs) yq_l = y[s];
                    const double yq = __shfl_sync(0xffffffffu, yq_l, ql);
                    const double sgd = (yq >= 0.0) ? delta : -delta;
                    const double beta = __drcp_rn(d2 + fabs(yq) * delta);
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
                    ++st.n_add;
                    __syncwarp();
                    PHASE(7);
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
                    symv_to_rowv(S, S.colk[1], nullptr, nV);         // rowv = H k   (barrier inside)
                    double kw = 0.0;
                    if (lane < RPW && row0 + lane < nV) kw = S.colk[1][row0 + lane] * S.rowv[row0 + lane];
                    kw = warp_sum_d(kw);
                    double rp[CS], kHk;
                    matvec_T(S, m, ybuf, S.rowv, kw, rp, kHk);       // rp_j = M[:,j]' (H k), kHk piggy-backed
                    const double ik = 1.0 / kHk;
                    const double rs = rsqrt(kHk);
                    const int q1 = q - 1, q1s = q1 >> 5, q1l = q1 & 31;
#pragma unroll
                    for (int r = 0; r < RPW; ++r) {
                        const double kr = S.colk[1][row0 + r];
#pragma unroll
                        for (int s = 0; s < CS; ++s) {
                            const int j = lane + 32 * s;
                            if (j < q && j != l) m[r][s] -= kr * (rp[s] * ik);
                        }
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
                    ++st.n_drop;
                    PHASE(8);
                }
            }
            if (failed) break;
        }
        __syncthreads();
        PHASE_FLUSH;
        return st;
    }

    // 1/2 x'Hx + g'x with the packed H (caller subtracts any regularisation); block-uniform result
    __device__ static double objective(SM& S, int nV, const double* diag_fix) {
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, row0 = warp * RPW;
        symv_to_rowv(S, S.x, nullptr, nV);
        double acc = 0.0;
        if (lane < RPW) {
            const int i = row0 + lane;
            if (i < nV) {
                double hx = S.rowv[i];
                if (diag_fix) hx -= diag_fix[i] * S.x[i];
                acc = S.x[i] * (0.5 * hx + S.g[i]);
            }
        }
        acc = warp_sum_d(acc);
        if (lane == 0) S.red_val[0][warp] = acc;
        __syncthreads();
        double f = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) f += S.red_val[0][w];
        __syncthreads();
        return f;
    }
};

}  // namespace fsae
