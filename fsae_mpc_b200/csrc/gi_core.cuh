// Register-tiled dual active-set core shared by the fused LTV-MPC kernels and the dense
// qpOASES drop-in kernel.
//
//   min 1/2 x'Hx + g'x   s.t. "slots" (variable bounds and general rows), each two-sided
//
// Goldfarb-Idnani (Math. Prog. 27, 1983) in operator form: the nV x nV matrix M = [K1 | J2]
//   J2 : H-orthonormal basis of the null space of the working set      (J2' H J2 = I, N' J2 = 0)
//   K1 : multiplier operator H^-1 N (N' H^-1 N)^-1                       (N' K1 = I, J2' H K1 = 0)
// lives in REGISTERS (GiTile; long horizons keep some column slots in shared memory): warp w owns rows
// w*RPW .. w*RPW+RPW-1, lane l owns columns l, l+32, ... (an RPW x CS tile per thread).  One iteration:
//   exact arg-min of the violations with two 32-bit REDUX per warp on an order-preserving key,
//   y = M'n (one cross-warp sum through smem; a normal with at most three entries -- SpN: variable bounds, rows on
//   integrator coordinates -- is a combination of rows of M that their owner warps publish instead: no dense
//   normal, no product),
//   z = J2 y2 (in-warp reduce-scatter; z_i stays in a register of the lane it ends up in, which alone updates x_i
//   and computes the row scalars k_i, w_i of the add), step lengths and the add/drop decision computed
//   redundantly by every warp, then either
//   add : Householder on J2 + rank-1 on K1 (one FMA per element outside the column slot of q), or
//   drop: K1 += k r'^T with r' = -K1' H k / k'Hk, the two column moves through shared memory by their owner lanes.
// Termination: a refresh (Newton step on the active manifold, multipliers from stationarity) after
// every run of partial steps; columns whose recomputed multiplier is negative are dropped.
// Template knobs (GiCfg): warps per problem NW, constraints taken per search KB (block adds),
// column slots in registers CSR.
// The problem-specific parts (how slots are evaluated and what their normals are) come from a policy
// object:  search(best, best_i),  normal_prepare(slot, side) + normal_entry(prep, i),  norm2(slot),
// is_unit(slot).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace fsae {

// Phase timers for kernel tuning (scripts/phase_profile.py builds a separate .so with
// -DFSAE_PROFILE; the product build compiles them out).
#ifdef FSAE_PROFILE
__device__ unsigned long long g_phase_cycles[16];
#define PHASE_DECL long long ph_t_ = clock64(); unsigned long long ph_acc_[16] = {0}
#define PHASE(i) do { const long long n_ = clock64(); ph_acc_[i] += (unsigned long long)(n_ - ph_t_); ph_t_ = n_; } while (0)
#define PHASE_FLUSH do { if (threadIdx.x == 0) for (int i_ = 0; i_ < 16; ++i_) atomicAdd(&g_phase_cycles[i_], ph_acc_[i_]); } while (0)
#define PHASE_COUNT(i) do { ++ph_acc_[i]; } while (0)
// stage timers of the enclosing kernel (tid 0): g_stage_cycles[i] += cycles since the previous mark
__device__ unsigned long long g_stage_cycles[16];
#define STAGE_DECL long long sg_t_ = clock64()
#define STAGE(i) do { if (threadIdx.x == 0) { const long long n_ = clock64(); atomicAdd(&g_stage_cycles[i], (unsigned long long)(n_ - sg_t_)); sg_t_ = n_; } } while (0)
#define DSTAGE_DECL long long ds_t_ = clock64()
#define DSTAGE(i) do { if (threadIdx.x == 0) { const long long n_ = clock64(); atomicAdd(&g_stage_cycles[12 + (i)], (unsigned long long)(n_ - ds_t_)); ds_t_ = n_; } } while (0)
#define RSTAGE_DECL long long rs_t_ = clock64()
#define RSTAGE(i) do { if (lane == 0) { const long long n_ = clock64(); atomicAdd(&g_stage_cycles[9 + (i)], (unsigned long long)(n_ - rs_t_)); rs_t_ = n_; } } while (0)
#else
#define RSTAGE_DECL
#define RSTAGE(i)
#define DSTAGE_DECL
#define DSTAGE(i)
#define STAGE_DECL
#define STAGE(i)
#define PHASE_DECL
#define PHASE(i)
#define PHASE_COUNT(i)
#define PHASE_FLUSH
#endif

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}


// ---- exact warp arg-min of doubles with two 32-bit REDUX ops -------------------------------
// dkey maps a double to an unsigned 64-bit key with the same order (negative < positive).
__device__ __forceinline__ unsigned long long dkey(double v) {
    const long long b = __double_as_longlong(v);
    return b < 0 ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dkey_inv(unsigned long long k) {
    const long long b = (k & 0x8000000000000000ull) ? (long long)(k & 0x7fffffffffffffffull) : (long long)~k;
    return __longlong_as_double(b);
}
constexpr unsigned long long DKEY_NONE = 0x8000000000000000ull;      // dkey(+0.0): "no candidate"
// lowest lane holding the minimum key; kmin = that key (identical in every lane)
__device__ __forceinline__ int warp_argmin_key(unsigned long long key, unsigned long long& kmin) {
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_min_sync(0xffffffffu, (hi == mh) ? lo : 0xffffffffu);
    const unsigned who = __ballot_sync(0xffffffffu, hi == mh && lo == ml);
    kmin = ((unsigned long long)mh << 32) | ml;
    return __ffs(who) - 1;
}

// ---- in-warp reduce-scatter of RH values (RH = 8, 16 or 32): afterwards lane l (and the lanes that
// differ from it only in the low log2(32/RH) bits) holds the warp-wide sum of entry l >> log2(32/RH).
template <int RH>
__device__ __forceinline__ double warp_reduce_scatter(double (&v)[RH]) {
    static_assert(RH == 8 || RH == 16 || RH == 32, "reduce-scatter width");
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int h = RH / 2, bit = 16; h >= 1; h >>= 1, bit >>= 1) {
        const bool hi = lane & bit;
#pragma unroll
        for (int i = 0; i < h; ++i) {
            const double send = hi ? v[i] : v[i + h];
            const double keep = hi ? v[i + h] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
    }
#pragma unroll
    for (int bit = 16 / RH; bit >= 1; bit >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], bit);
    return v[0];
}

// A constraint normal with at most three entries (rows ascending), cnt = 0: dense.  Scalars, not arrays: the
// entries must stay in registers (an array that is ever indexed at run time lives in local memory).
struct SpN {
    int cnt, i0, i1, i2;
    double c0, c1, c2;
    __device__ __forceinline__ int idx(int e) const { return e == 0 ? i0 : (e == 1 ? i1 : i2); }
    __device__ __forceinline__ double cf(int e) const { return e == 0 ? c0 : (e == 1 ? c1 : c2); }
};
__device__ __forceinline__ SpN spn_none() { return SpN{0, 0, 0, 0, 0.0, 0.0, 0.0}; }

// Tile geometry for at most NVMAX variables and NW warps per CTA.
template <int NVMAX, int NW_ = 8, int KB_ = 1, int CSR_ = -1>
struct GiCfg {
    static constexpr int NW = NW_, NT = 32 * NW_;
    static constexpr int KB = KB_;                           // constraints projected per search (block size)
    static constexpr int YB = 2;                             // ypart buffers (double-buffered)
    static_assert(KB_ >= 1 && NW_ * KB_ <= 32, "block selection reduces NW*KB candidates in one warp");
    static constexpr int RPW = (NVMAX + NW - 1) / NW;      // rows per warp
    static constexpr int RP = RPW * NW;                      // padded rows
    static constexpr int CS = (NVMAX + 32) / 32;             // column slots per lane (one spare column)
    static constexpr int CP = CS * 32;                       // padded columns
    static constexpr int SP = (NVMAX > 96) ? 8 : 3;          // partial sums per row of a packed symv (long horizons: H in L2); >= 2
    static constexpr bool TWO_LEVEL_SUM = NW_ >= 6;          // cross-warp sum of M'v: slice per warp + second barrier
    static constexpr int CSR = (CSR_ < 0 || CSR_ > CS) ? CS : CSR_;   // column slots held in registers ...
    static constexpr int CSS = CS - CSR;                     // ... and in shared memory (long horizons)
    static constexpr int RH = (RPW <= 8) ? 8 : (RPW <= 16 ? 16 : 32);   // reduce-scatter width
    static constexpr int HP = NVMAX * (NVMAX + 1) / 2;       // packed lower triangle
    static_assert(RPW <= 32, "reduce-scatter network supports up to 32 rows per warp");
    static_assert(NVMAX < CP, "need one spare padded column for the piggy-backed scalar");
    __host__ __device__ static constexpr int hp(int i, int j) { return i * (i + 1) / 2 + j; }   // i >= j
};

// Operator tile of one thread: RPW rows x CS column slots.  The first CSR slots live in registers, the
// remaining CSS = CS - CSR in shared memory (long horizons: the operator of nV = 161 does not fit the
// register file of one SM).  Slot indices are compile-time constants after unrolling, so the choice
// folds away; the shared part is laid out [warp][row][slot][lane] (conflict-free).
template <class G>
struct GiTile {
    static constexpr int RPW = G::RPW, CSR = G::CSR, CSS = G::CSS;
    double reg[RPW][CSR];
    double* sm;            // this thread's first shared-memory element
    static constexpr int SM_DOUBLES = G::NW * RPW * (CSS > 0 ? CSS : 0) * 32;
    __device__ __forceinline__ void attach(double* base) {
        sm = base + (size_t)(threadIdx.x >> 5) * RPW * CSS * 32 + (threadIdx.x & 31);
    }
    __device__ __forceinline__ double& operator()(int r, int s) {
        if (CSS == 0 || s < CSR) return reg[r][s < CSR ? s : 0];
        return sm[(r * CSS + (s - CSR)) * 32];
    }
    __device__ __forceinline__ const double& operator()(int r, int s) const {
        if (CSS == 0 || s < CSR) return reg[r][s < CSR ? s : 0];
        return sm[(r * CSS + (s - CSR)) * 32];
    }
};

// Shared-memory working set of the core.
// WSP > 0: the first WSP variables are held in INTEGRATOR COORDINATES w = T u (two interleaved control channels,
// w_{2k+c} = dt * sum_{i<=k} u_{2i+c}; see fused_v2.cuh) while the packed H stays in u coordinates: every H v
// becomes T^-T H T^-1 v, two local difference stencils around the same packed product.
template <class G, int NSLOT, bool HPG = false, int WSP_ = 0>
struct GiSm {
    static constexpr int WSP = WSP_;
    double idt;                        // 1 / dt (WSP > 0)
    alignas(16) double x[G::RP];
    double g[G::RP];
    double Hp[HPG ? 2 : G::HP];        // packed lower triangle of H (drops, refresh, fval) ...
    double* hpg;                       // ... or, for long horizons, the FULL symmetric H in a per-problem global (L2-resident) slab
    static constexpr bool HFULL = HPG;
    __device__ __forceinline__ double* hp() { return HPG ? hpg : Hp; }
    double ypart[G::YB][G::NW][G::CP]; // cross-warp partial sums of M'v (double-buffered; one buffer per block member)
    double colk[2][G::RP];             // column broadcast (factorisation / column q / drop column)
    double rowv[G::RP];                // per-warp row vector scratch (gradient / H k)
    double nvec[G::KB][G::RP];         // per-warp rows of the normals of the block's constraints
    double zrow[G::RP];                // per-warp reduced z
    double ysum[G::TWO_LEVEL_SUM ? G::CP : 1];   // published cross-warp sums of M'v (two-level variant)
    double xs0[G::RP];                 // x as the block's search saw it (own rows per warp)
    alignas(16) double wpart[G::SP][G::RP];   // symv partials; between symv's also the (k_i, w_i) pairs of the add update
    double dvec[G::RP];                // LDL' pivots
    double red_val[2][G::NW];
    unsigned long long red_key[2][G::NW * G::KB];   // per-warp top-KB violations (dkey) ...
    int red_idx[2][G::NW * G::KB];                  // ... and their slot*2 + side codes
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
    __device__ __forceinline__ static void matvec_T(SM& S, const GiTile<G>& m, int& ybuf,
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
            for (int s = 0; s < CS; ++s) yp[s] += m(r, s) * v;
        }
        if (lane == 31) yp[CS - 1] = extra;
#pragma unroll
        for (int s = 0; s < CS; ++s) S.ypart[ybuf][warp][lane + 32 * s] = yp[s];
        __syncthreads();
        if constexpr (G::TWO_LEVEL_SUM) {
            // many warps (long horizons): every warp sums a slice of the columns once and publishes it,
            // instead of every warp summing every column (NW * CS loads per thread)
            constexpr int CPW = (G::CP + NW - 1) / NW;
            static_assert(CPW <= 32, "one lane per column of the slice");
            const int col = warp * CPW + lane;
            if (lane < CPW && col < G::CP) {
                double acc = 0.0;
#pragma unroll
                for (int w = 0; w < NW; ++w) acc += S.ypart[ybuf][w][col];
                S.ysum[col] = acc;
            }
            __syncthreads();
#pragma unroll
            for (int s = 0; s < CS; ++s) y[s] = S.ysum[lane + 32 * s];
        } else {
#pragma unroll
            for (int s = 0; s < CS; ++s) {
                double acc = 0.0;
#pragma unroll
                for (int w = 0; w < NW; ++w) acc += S.ypart[ybuf][w][lane + 32 * s];
                y[s] = acc;
            }
        }
        ybuf ^= 1;
        extra_sum = __shfl_sync(0xffffffffu, y[CS - 1], 31);
        if (lane == 31) y[CS - 1] = 0.0;
    }

    // z = sum_{q0 <= j < nV} M[:, j] y_j for this warp's rows -> S.zrow, visible after __syncwarp
    __device__ __forceinline__ static void matvec_N(SM& S, const GiTile<G>& m, const double (&y)[CS],
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
            for (int r = 0; r < RPW; ++r) zp[r] += m(r, s) * yj;
        }
        const double zr = warp_reduce_scatter<RH>(zp);
        constexpr int SH = (RH == 32) ? 0 : (RH == 16 ? 1 : 2);
        const int rr = lane >> SH;
        const bool writer = (lane & ((1 << SH) - 1)) == 0;
        if (writer && rr < RPW) S.zrow[row0 + rr] = zr;
        __syncwarp();
    }

    // The same product for the add iteration: the sum of row rr = lane >> SH stays in a REGISTER of the lanes that
    // the reduce-scatter leaves it in (`holds`: the first lane of each group, rr < RPW) -- x, k and w of that row are
    // computed by that lane, so z never goes through shared memory.
    static constexpr int ZSH = (RH == 32) ? 0 : (RH == 16 ? 1 : 2);
    __device__ __forceinline__ static double matvec_N_reg(const GiTile<G>& m, const double (&y)[CS], int q0, int nV,
                                                          int& rr, bool& holds) {
        const int lane = threadIdx.x & 31;
        double zp[RH];
#pragma unroll
        for (int r = 0; r < RH; ++r) zp[r] = 0.0;
#pragma unroll
        for (int s = 0; s < CS; ++s) {
            const int j = lane + 32 * s;
            const double yj = (j >= q0 && j < nV) ? y[s] : 0.0;
#pragma unroll
            for (int r = 0; r < RPW; ++r) zp[r] += m(r, s) * yj;
        }
        const double zr = warp_reduce_scatter<RH>(zp);
        rr = lane >> ZSH;
        holds = ((lane & ((1 << ZSH) - 1)) == 0) && rr < RPW;
        return zr;
    }

    // rowv = Hp * v (+ addv) for this warp's rows; v full length in shared; one barrier inside
    __device__ __forceinline__ static void symv_to_rowv(SM& S, const double* v, const double* addv, int nV) {
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, row0 = warp * RPW;
        constexpr int SP = G::SP;
        const int CH = (nV + SP - 1) / SP;
        if constexpr (SM::WSP > 0) {
            // T^-1 v: back to control increments (difference along each channel, / dt)
            for (int i = tid; i < nV; i += NT) S.dvec[i] = (i < SM::WSP) ? (v[i] - (i >= 2 ? v[i - 2] : 0.0)) * S.idt : v[i];
            __syncthreads();
            v = S.dvec;
        }
        for (int t = tid; t < SP * nV; t += NT) {
            const int pt = t / nV, i = t - pt * nV;
            const int j0 = pt * CH, j1 = (j0 + CH < nV) ? j0 + CH : nV;
            double acc = 0.0;
            if (!SM::HFULL) {
                for (int j = j0; j < j1; ++j) acc += ((j <= i) ? S.hp()[G::hp(i, j)] : S.hp()[G::hp(j, i)]) * v[j];
            } else {
                // full symmetric H in the L2 slab, read by columns (H[j][i] = H[i][j]): consecutive threads read
                // consecutive addresses; four independent loads in flight
                const double* H = S.hp() + i;
                double a[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) a[u] = 0.0;
                for (int j = j0; j < j1; j += 8) {
                    double h[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) h[u] = (j + u < j1) ? H[(size_t)(j + u) * nV] : 0.0;     // eight L2 loads in flight
#pragma unroll
                    for (int u = 0; u < 8; ++u) a[u] = fma(h[u], (j + u < j1) ? v[j + u] : 0.0, a[u]);
                }
                acc = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
            }
            S.wpart[pt][i] = acc;
        }
        __syncthreads();
        if (lane < RPW) {
            const int i = row0 + lane;
            if (i < nV) {
                double acc = S.wpart[0][i];
#pragma unroll
                for (int pt = 1; pt < SP; ++pt) acc += S.wpart[pt][i];
                if constexpr (SM::WSP > 0) {
                    if (i < SM::WSP) {             // T^-T: difference towards the later step, / dt
                        double nxt = 0.0;
                        if (i + 2 < SM::WSP) {
                            nxt = S.wpart[0][i + 2];
#pragma unroll
                            for (int pt = 1; pt < SP; ++pt) nxt += S.wpart[pt][i + 2];
                        }
                        acc = (acc - nxt) * S.idt;
                    }
                }
                S.rowv[i] = acc + (addv ? addv[i] : 0.0);
            }
        }
        __syncwarp();
    }

    // Symmetric elimination of the leading nC x nC block held in the tiles (the "curved"
    // variables) with the column operations accumulated in place, then the layout
    //   M = [ e_{nC}, .., e_{nC+ns-1} | J ],  J = L^-T (J J' = H_cc^-1), flat variables last.
    // The ns flat (zero-curvature) variables start with one bound in the working set:
    // q = ns, lam_j = |g_flat_j|.  One column broadcast + one barrier per elimination step.
    // Returns false (uniformly) if a pivot is not positive.
    __device__ static bool factor_and_layout(SM& S, GiTile<G>& m, double (&lam)[CS], int& q,
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
                        for (int r = 0; r < RPW; ++r) S.colk[buf][row0 + r] = m(r, s);
                    }
            }
            __syncthreads();
            const double piv = S.colk[buf][k];
            ok = ok && (piv > 0.0);
            const double rp = __drcp_rn(piv);      // correctly rounded, same value as 1.0 / piv
            if (tid == 0) S.dvec[k] = piv;
            double lj[CS];
#pragma unroll
            for (int s = 0; s < CS; ++s) {
                const int j = lane + 32 * s;
                lj[s] = (j > k && j < nC) ? S.colk[buf][j] * rp : 0.0;     // symmetric: W[k][j] = W[j][k]
            }
            // column slots that lie entirely at or left of the pivot have lj = 0: skipped (warp-uniform)
            const int s0 = (k + 1) >> 5;                // first slot that still has a column right of the pivot
#pragma unroll
            for (int sb = 0; sb < CS; ++sb) {
                if (s0 == sb) {
#pragma unroll
                    for (int r = 0; r < RPW; ++r) {
                        const double vr = S.colk[buf][row0 + r];
#pragma unroll
                        for (int s = sb; s < CS; ++s) m(r, s) = fma(-vr, lj[s], m(r, s));
                    }
                }
            }
            const int kr = k - row0;                    // warp-uniform: does this warp own the pivot row?
            if (kr >= 0 && kr < RPW) {
#pragma unroll
                for (int r = 0; r < RPW; ++r)
                    if (r == kr) {
#pragma unroll
                        for (int s = 0; s < CS; ++s) {
                            const int j = lane + 32 * s;
                            if (j > k && j < nC) m(r, s) = -lj[s];
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
                if (j < nC && i < nC) v = (i < j) ? m(r, s) * sc : (i == j ? sc : 0.0);
                m(r, s) = v;
            }
        }
        // rotate the J columns right by ns lanes (K1 first): column j comes from column j - ns
        {
            double t[RPW][CS];
#pragma unroll
            for (int s = 0; s < CS; ++s) {
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    const double same = __shfl_sync(0xffffffffu, m(r, s), (lane - ns) & 31);
                    const double prev = (s > 0) ? __shfl_sync(0xffffffffu, m(r, s > 0 ? s - 1 : 0), (lane - ns) & 31) : 0.0;
                    t[r][s] = (lane >= ns) ? same : prev;
                }
            }
#pragma unroll
            for (int s = 0; s < CS; ++s) {
                const int j = lane + 32 * s;
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    const int i = row0 + r;
                    // K1 column of a flat variable that starts on its bound: N'K1 = I with the normal
                    // +e_i at a lower bound (x_i >= lb) and -e_i at an upper bound (-x_i >= -ub)
                    const double sgn = (j < ns && (S.act[j < ns ? j : 0] & 1)) ? -1.0 : 1.0;
                    m(r, s) = (j < ns) ? ((i == nC + j) ? sgn : 0.0) : (j < nV ? t[r][s] : 0.0);
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
    __device__ static void initial_point(SM& S, const GiTile<G>& m, int& ybuf, int q, int nC, int nV) {
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
    //
    // Block selection: one search yields the KB most violated constraint sides.  The first goes
    // through the full Goldfarb-Idnani step (partial steps / drops included); the others
    // "piggy-back" on the same search: their normals are already in shared memory, their
    // current violation follows from the x the search saw, s_c += n_c'(x - x_search), which rides
    // on the spare column of the projection pass y = M'n_c, so no second evaluation of the
    // constraints is needed.  A piggy-backed candidate that is no longer violated is skipped; one
    // that would need a partial step ends the block (the next search finds it again).  Between
    // the adds of a block x, z and the tile updates are warp-local: the only block barrier of a
    // piggy-backed add is the one inside the projection.
    template <class Prob>
    __device__ static GiStats solve(const Prob& prob, SM& S, GiTile<G>& m, double (&lam)[CS], int& q,
                                    int& ybuf, int nV, double tol, int max_iter) {
        constexpr int KB = G::KB;
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, row0 = warp * RPW;
        GiStats st = {0, GI_EXIT_SOLVED, 0, 0, 0};
        int rbuf = 0, drops_at_refresh = 0;
        bool refresh_failed = false;
        unsigned long long my_key = DKEY_NONE;      // candidate reuse: this warp's most violated slot of the last full search
        int my_code = 0x7fffffff;
        bool have_cand = false;
        PHASE_DECL;
        // drop working-set column l: K1 <- K1 + k r'^T with r' = -K1' H k / k'Hk (k = column l), the freed
        // direction k / sqrt(k'Hk) joins J2 as column q-1, column q-1 of K1 moves into slot l
        auto drop_column = [&](int l) {
            DSTAGE_DECL;
            const int ls = l >> 5, ll = l & 31;
            if (lane == ll) {
#pragma unroll
                for (int s = 0; s < CS; ++s)
                    if (s == ls) {
#pragma unroll
                        for (int r = 0; r < RPW; ++r) S.colk[1][row0 + r] = m(r, s);
                    }
            }
            __syncthreads();                                 // k = M[:, l] visible block-wide
            DSTAGE(0);
            symv_to_rowv(S, S.colk[1], nullptr, nV);         // rowv = H k   (barrier inside)
            DSTAGE(1);
            double kw = 0.0;
            if (lane < RPW && row0 + lane < nV) kw = S.colk[1][row0 + lane] * S.rowv[row0 + lane];
            kw = warp_sum_d(kw);
            double rp[CS], kHk;
            matvec_T(S, m, ybuf, S.rowv, kw, rp, kHk);       // rp_j = M[:,j]' (H k), kHk piggy-backed
            DSTAGE(2);
            const double ik = 1.0 / kHk;
            const double rs = rsqrt(kHk);
            const int q1 = q - 1, q1s = q1 >> 5, q1l = q1 & 31;
            // K1 columns j < q, j != l:  m - k (r'_j);  the masked, scaled r' once per lane
            double rpm[CS];
#pragma unroll
            for (int s = 0; s < CS; ++s) {
                const int j = lane + 32 * s;
                rpm[s] = (j < q && j != l) ? rp[s] * ik : 0.0;
            }
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const double kr = S.colk[1][row0 + r];
#pragma unroll
                for (int s = 0; s < CS; ++s) m(r, s) = fma(-kr, rpm[s], m(r, s));
            }
            // column moves through shared memory (this warp's rows only): the updated column q-1 goes into slot l, the
            // freed direction k / sqrt(k'Hk) becomes column q-1.  Only the two owner lanes touch their registers.
            const bool mv = (l != q1);
            if (mv && lane == q1l) {
#pragma unroll
                for (int s = 0; s < CS; ++s)
                    if (s == q1s) {
#pragma unroll
                        for (int r = 0; r < RPW; ++r) S.colk[0][row0 + r] = m(r, s);
                    }
            }
            __syncwarp();
            if (mv && lane == ll) {
#pragma unroll
                for (int s = 0; s < CS; ++s)
                    if (s == ls) {
#pragma unroll
                        for (int r = 0; r < RPW; ++r) m(r, s) = S.colk[0][row0 + r];
                    }
            }
            if (lane == q1l) {
#pragma unroll
                for (int s = 0; s < CS; ++s)
                    if (s == q1s) {
#pragma unroll
                        for (int r = 0; r < RPW; ++r) m(r, s) = S.colk[1][row0 + r] * rs;
                    }
            }
            __syncwarp();
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
            DSTAGE(3);
        };
        while (true) {
            // P1: the KB most violated inactive constraint sides (policy evaluates its slots; one
            // candidate per thread).
            // Candidate reuse (Prob::REUSE, KB = 1): a full search leaves every warp with its own most violated
            // slot.  After the global winner has been added, the other warps' candidates are usually still
            // violated; each warp re-evaluates ITS candidate exactly against the current x (one short
            // warp-cooperative dot product instead of all slots) and the most violated of them is taken.  Any
            // violated constraint is a valid pivot of the dual method, and the solve only ends after a FULL search
            // finds none, so the minimiser (unique) is the same; only the pivot order differs.
            double cviol[KB];
            int ccode[KB];
            bool picked = false;
            if constexpr (Prob::REUSE && KB == 1) {
                if (have_cand) {
                    PHASE(0);
                    double v = 0.0;
                    if (my_key != DKEY_NONE && S.status[my_code >> 1] == 0) v = prob.eval_code(my_code);
                    const unsigned long long key = (v < 0.0) ? dkey(v) : DKEY_NONE;
                    if (lane == 0) { S.red_key[rbuf][warp] = key; S.red_idx[rbuf][warp] = my_code; }
                    my_key = key;
                    PHASE(9);
                    __syncthreads();
                    unsigned long long ck = DKEY_NONE, km;
                    int ci = 0x7fffffff;
                    if (lane < NW) { ck = S.red_key[rbuf][lane]; ci = S.red_idx[rbuf][lane]; }
                    rbuf ^= 1;
                    const int wl = warp_argmin_key(ck, km);
                    ccode[0] = __shfl_sync(0xffffffffu, ci, wl);
                    cviol[0] = dkey_inv(km);
                    if (cviol[0] < -tol) {
                        picked = true;
                        if (warp == wl) my_key = DKEY_NONE;          // consumed
                        PHASE_COUNT(13);
                    } else {
                        have_cand = false;
                    }
                    PHASE(2);
                }
            }
            if (!picked) {
                double best = 0.0;
                int best_i = 0x7fffffff;
                PHASE(0);
                PHASE_COUNT(12);
                prob.search(best, best_i);
                PHASE(1);
                unsigned long long key = dkey(best);
#pragma unroll
                for (int c = 0; c < KB; ++c) {
                    unsigned long long km;
                    const int wl = warp_argmin_key(key, km);
                    const int wi = __shfl_sync(0xffffffffu, best_i, wl);
                    if (lane == 0) { S.red_key[rbuf][warp * KB + c] = km; S.red_idx[rbuf][warp * KB + c] = wi; }
                    if (Prob::REUSE && KB == 1) { my_key = km; my_code = wi; }
                    if (lane == wl) key = DKEY_NONE;
                }
                __syncthreads();
                unsigned long long ck = DKEY_NONE;
                int ci = 0x7fffffff;
                if (lane < NW * KB) { ck = S.red_key[rbuf][lane]; ci = S.red_idx[rbuf][lane]; }
                rbuf ^= 1;
#pragma unroll
                for (int c = 0; c < KB; ++c) {
                    unsigned long long km;
                    const int wl = warp_argmin_key(ck, km);
                    ccode[c] = __shfl_sync(0xffffffffu, ci, wl);
                    cviol[c] = dkey_inv(km);
                    if (Prob::REUSE && KB == 1 && warp == wl) my_key = DKEY_NONE;      // the winner is being added
                    if (lane == wl) ck = DKEY_NONE;
                }
                have_cand = Prob::REUSE && KB == 1;
            }
            PHASE(2);

            if (!(cviol[0] < -tol)) {
                // the refresh repairs what chains of partial steps leave behind; a run of pure
                // full steps keeps x the exact working-set minimiser (to round-off)
                if (st.n_drop == drops_at_refresh) break;
                // refresh: Newton step on the active manifold + multipliers from stationarity
                // (again whenever further partial steps followed the previous refresh)
                ++st.n_refresh;
                drops_at_refresh = st.n_drop;
                // Round-off in long chains of partial steps (multipliers of size R_soft ~ 1e8 next to
                // multipliers of size 1) can leave a constraint in the working set whose true multiplier
                // is negative.  The multipliers recomputed from stationarity show it: such a column is
                // dropped and the step repeated (then the loop goes on from a dual-feasible point).
                for (int pass = 0;; ++pass) {
                    symv_to_rowv(S, S.x, S.g, nV);                       // rowv = H x + g
                    double y[CS], dummy;
                    matvec_T(S, m, ybuf, S.rowv, 0.0, y, dummy);
                    matvec_N(S, m, y, q, nV);
                    if (lane < RPW) {
                        const int i = row0 + lane;
                        if (i < nV) S.x[i] -= S.zrow[i];
                    }
                    double ymin = 0.0, ymax = 0.0;
                    int lmin = -1;
#pragma unroll
                    for (int s = 0; s < CS; ++s) {
                        const int j = lane + 32 * s;
                        if (j < q) {
                            lam[s] = fmax(y[s], 0.0);
                            ymax = fmax(ymax, fabs(y[s]));
                            if (y[s] < ymin) { ymin = y[s]; lmin = j; }
                        }
                    }
                    {
                        unsigned long long km;
                        const int wl = warp_argmin_key(dkey(ymin), km);
                        lmin = __shfl_sync(0xffffffffu, lmin, wl);
                        ymin = dkey_inv(km);
                        (void)warp_argmin_key(dkey(-ymax), km);
                        ymax = -dkey_inv(km);
                    }
                    __syncthreads();                   // x complete
                    if (!(ymin < -1e-10 * (1.0 + ymax))) break;
                    // a multiplier is still negative: the point is primal feasible but not optimal.  Out of
                    // passes / iterations that is qpOASES's "maximum number of iterations" (exitflag 1), not success.
                    if (pass >= 8 || ++st.iters > max_iter) { st.exitflag = GI_EXIT_MAXITER; refresh_failed = true; break; }
                    drop_column(lmin);
                    drops_at_refresh = st.n_drop;
                }
                if (refresh_failed) break;
                continue;
            }
            int nleft = 0;                          // candidates (sorted by violation: a prefix is valid)
#pragma unroll
            for (int c = 0; c < KB; ++c) nleft += (cviol[c] < -tol) ? 1 : 0;
            if (q + nleft > nV) nleft = nV - q > 0 ? nV - q : 1;

            // Sparse normals (KB = 1): a normal with at most three entries (a variable bound, a row that touches only
            // integrator coordinates and a slack) needs neither its dense vector nor the product M'n -- y is a
            // combination of at most three ROWS of M, which their owners publish.
            SpN spn = spn_none();
            if constexpr (KB == 1) spn = prob.sparse_normal(ccode[0] >> 1, (ccode[0] & 1) ? +1 : -1);
            const int scnt = spn.cnt;
            // P2: this warp's entries of the block's normals; the x the search saw (own rows)
#pragma unroll
            for (int c = 0; c < KB; ++c) {
                if (c < nleft && scnt == 0) {
                    const auto prep = prob.normal_prepare(ccode[c] >> 1, (ccode[c] & 1) ? +1 : -1);   // warp-uniform part
                    if (lane < RPW) {
                        const int i = row0 + lane;
                        S.nvec[c][i] = (i < nV) ? prob.normal_entry(prep, i) : 0.0;
                    }
                }
            }
            if (KB > 1 && nleft > 1 && lane < RPW) S.xs0[row0 + lane] = S.x[row0 + lane];
            __syncwarp();
            PHASE(3);

            int cidx = 0;                           // which of the block's normals is being added
            int pslot = ccode[0] >> 1, pside = (ccode[0] & 1) ? +1 : -1;
            double sp = cviol[0];                   // n'x - b  (< 0)
            double lam_p = 0.0;
            double nn = prob.norm2(pslot);
            bool piggy = false, dropped = false, failed = false;
            while (true) {
                if (++st.iters > max_iter) { st.exitflag = GI_EXIT_MAXITER; failed = true; break; }
                // P3: y = M'n.  A variable bound has n = +-e_p, so y is +-(row p of M): the owning warp
                // publishes its row, nobody multiplies or sums partials.  The spare column carries
                // n'(x - x_search) for a piggy-backed candidate.
                double y[CS], dsum = 0.0;
                if (KB == 1 && scnt > 0) {
                    double yp[CS];
#pragma unroll
                    for (int s = 0; s < CS; ++s) yp[s] = 0.0;
                    bool mine = false;
#pragma unroll
                    for (int e = 0; e < 3; ++e) {
                        const int pr = spn.idx(e) - row0;          // warp-uniform
                        if (e < scnt && pr >= 0 && pr < RPW) {
                            mine = true;
#pragma unroll
                            for (int r = 0; r < RPW; ++r)
                                if (r == pr) {
#pragma unroll
                                    for (int s = 0; s < CS; ++s) yp[s] = fma(spn.cf(e), m(r, s), yp[s]);
                                }
                        }
                    }
                    if (mine) {
#pragma unroll
                        for (int s = 0; s < CS; ++s) S.ypart[ybuf][warp][lane + 32 * s] = yp[s];
                    }
                    __syncthreads();
#pragma unroll
                    for (int s = 0; s < CS; ++s) y[s] = 0.0;
                    int lastw = -1;
#pragma unroll
                    for (int e = 0; e < 3; ++e) {                  // entries ascend: the owners of equal rows are adjacent
                        const int w = spn.idx(e) / RPW;
                        if (e < scnt && w != lastw) {
#pragma unroll
                            for (int s = 0; s < CS; ++s) y[s] += S.ypart[ybuf][w][lane + 32 * s];
                            lastw = w;
                        }
                    }
                    ybuf ^= 1;
                } else if (prob.is_unit(pslot)) {
                    const int pr = pslot - row0;               // warp-uniform
                    if (pr >= 0 && pr < RPW) {
#pragma unroll
                        for (int r = 0; r < RPW; ++r)
                            if (r == pr) {
#pragma unroll
                                for (int s = 0; s < CS; ++s) S.ypart[ybuf][0][lane + 32 * s] = m(r, s);
                            }
                        if (KB > 1 && piggy && lane == 31) S.ypart[ybuf][0][G::CP - 1] = S.x[pslot] - S.xs0[pslot];
                    }
                    __syncthreads();
                    const double sgn = pside < 0 ? 1.0 : -1.0;
#pragma unroll
                    for (int s = 0; s < CS; ++s) y[s] = sgn * S.ypart[ybuf][0][lane + 32 * s];
                    if (KB > 1 && piggy) {
                        dsum = __shfl_sync(0xffffffffu, y[CS - 1], 31);
                        if (lane == 31) y[CS - 1] = 0.0;
                    }
                    ybuf ^= 1;
                } else {
                    double extra = 0.0;
                    if (KB > 1 && piggy) {
                        if (lane < RPW) extra = S.nvec[cidx][row0 + lane] * (S.x[row0 + lane] - S.xs0[row0 + lane]);
                        extra = warp_sum_d(extra);
                    }
                    matvec_T(S, m, ybuf, S.nvec[cidx], extra, y, dsum);
                }
                // z = J2 y2 needs only y: issued here so that its FMAs and reduce-scatter shuffles interleave
                // with the reductions of the step-length phase (independent dependency chains)
                {
                    // column q (it leaves J2 if p is added): published now, read by the row lanes after the step lengths
                    const int qs0 = q >> 5, ql0 = q & 31;
                    if (lane == ql0) {
#pragma unroll
                        for (int s = 0; s < CS; ++s)
                            if (s == qs0) {
#pragma unroll
                                for (int r = 0; r < RPW; ++r) S.colk[0][row0 + r] = m(r, s);
                            }
                    }
                }
                int zrr;
                bool zholds;
                const double zr_mine = matvec_N_reg(m, y, q, nV, zrr, zholds);
                PHASE(5);
                bool more = true;                   // (piggy-backed only) keep going with this candidate
                if (KB > 1 && piggy) {
                    sp = cviol[0] + dsum;
                    if (!(sp < -tol)) { more = false; PHASE_COUNT(15); }
                }
                double d2 = 0.0, t1 = INFINITY, inv_d2 = 0.0, t = 0.0, sgd = 0.0, beta = 0.0;
                int l = -1;
                bool full = false, primal = false;
                if (more) {
                    // P4 (every warp, redundantly): step lengths
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
                        unsigned long long km;
                        const int wl = warp_argmin_key(dkey(t1), km);
                        l = __shfl_sync(0xffffffffu, l, wl);
                        t1 = dkey_inv(km);
                    }
                    const bool lin_dep = !(d2 > 1e-13 * fmax(1.0, nn));
                    const double rs2 = rsqrt(d2);
                    inv_d2 = rs2 * rs2;             // 2-3 ulp; only scales step lengths and the new column (one special-function chain less)
                    {
                        // Householder scalars of a possible add (rsqrt + reciprocal: ~150 cycles of dependent latency),
                        // issued here so that they overlap the primal step instead of delaying the tile update
                        const int qs = q >> 5, ql = q & 31;
                        double yq_l = 0.0;
#pragma unroll
                        for (int s = 0; s < CS; ++s)
                            if (s == qs) yq_l = y[s];
                        const double yqv = __shfl_sync(0xffffffffu, yq_l, ql);
                        const double delta = d2 * rs2;
                        sgd = (yqv >= 0.0) ? delta : -delta;
                        beta = __drcp_rn(d2 + fabs(yqv) * delta);
                    }
                    const double t2 = lin_dep ? INFINITY : (sp < 0.0 ? -sp * inv_d2 : 0.0);
                    full = (t2 <= t1);
                    if (piggy && !full) { --st.iters; PHASE_COUNT(14); break; }     // left to the next search
                    if (isinf(t1) && isinf(t2)) { st.exitflag = GI_EXIT_INFEASIBLE; failed = true; break; }
                    primal = !isinf(t2);
                    t = full ? t2 : t1;
                    PHASE(6);
                    // P5: x += t z  (this warp's rows only)
                    if (primal) {
                        if (zholds && row0 + zrr < nV) S.x[row0 + zrr] += t * zr_mine;
                        sp += t * d2;
                    }
                    if (full && tid == 0) {
                        // bookkeeping of the add; published by the barrier that ends the block
                        S.act[q] = pslot * 2 + (pside > 0 ? 1 : 0);
                        S.status[pslot] = (int8_t)pside;
                    }
#pragma unroll
                    for (int s = 0; s < CS; ++s) {
                        const int j = lane + 32 * s;
                        if (j < q) lam[s] -= t * y[s];
                    }
                    lam_p += t;
                    PHASE(7);
                } else {
                    --st.iters;
                }
                if (!more || full) {
                    if (more) {
                        // P6a: add p.  K1 <- K1 - k r', J2 <- J2 (I - beta v v'), column q <- k = z/d2
                        const int qs = q >> 5;
                        __syncwarp();                       // column q visible to the row lanes
                        if (zholds) {                       // k_i = z_i / d2,  w_i = beta (z_i + sgd M[i][q]): once per row
                            double2 kwv;
                            kwv.x = zr_mine * inv_d2;
                            kwv.y = (zr_mine + sgd * S.colk[0][row0 + zrr]) * beta;
                            reinterpret_cast<double2*>(&S.wpart[0][0])[row0 + zrr] = kwv;      // (k_i, w_i): wpart is free between symv's
                        }
                        __syncwarp();
                        // Column slots entirely left of q take  m - kr y,  slots entirely right of it
                        // m - wr y  (one FMA per element, warp-uniform choice); only the slot that holds
                        // column q mixes the three cases:  c*m - kr*ya - wr*yb  with (c, ya, yb) =
                        // (1, y, 0) for j < q, (0, -1, 0) for j == q, (1, 0, y) for j > q.
                        double ccq = 1.0, yaq = 0.0, ybq = 0.0;
#pragma unroll
                        for (int s = 0; s < CS; ++s) {
                            if (s == qs) {
                                const int j = lane + 32 * s;
                                ccq = (j == q) ? 0.0 : 1.0;
                                yaq = (j < q) ? y[s] : (j == q ? -1.0 : 0.0);
                                ybq = (j > q) ? y[s] : 0.0;
                            }
                        }
#pragma unroll
                        for (int qq = 0; qq < CS; ++qq) {
                            if (qs == qq) {                 // one specialised copy per position of the mixed slot
#pragma unroll
                                for (int r = 0; r < RPW; ++r) {
                                    const double2 kwv = reinterpret_cast<const double2*>(&S.wpart[0][0])[row0 + r];
                                    const double kr = kwv.x, wr = kwv.y;
#pragma unroll
                                    for (int s = 0; s < CS; ++s) {
                                        if (s < qq) m(r, s) = fma(-kr, y[s], m(r, s));
                                        else if (s > qq) m(r, s) = fma(-wr, y[s], m(r, s));
                                        else m(r, s) = fma(-wr, ybq, fma(-kr, yaq, ccq * m(r, s)));
                                    }
                                }
                            }
                        }
#pragma unroll
                        for (int s = 0; s < CS; ++s) {
                            const int j = lane + 32 * s;
                            if (j == q) lam[s] = lam_p;
                        }
                        ++q;
                        ++st.n_add;
                        __syncwarp();
                        PHASE(8);
                    }
                    // next candidate of the block
                    if (KB == 1 || dropped || nleft <= 1 || q >= nV) break;
#pragma unroll
                    for (int c = 0; c + 1 < KB; ++c) {
                        cviol[c] = cviol[c + 1];
                        ccode[c] = ccode[c + 1];
                    }
                    --nleft;
                    ++cidx;
                    pslot = ccode[0] >> 1;
                    pside = (ccode[0] & 1) ? +1 : -1;
                    lam_p = 0.0;
                    nn = prob.norm2(pslot);
                    piggy = true;
                    PHASE_COUNT(13);
                    PHASE(9);
                    continue;
                }
                // P6b: drop active constraint l (column l of K1)
                dropped = true;
                drop_column(l);
                PHASE(10);

            }
            if (failed) break;
            __syncthreads();                       // x, act, status of the whole block published
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
