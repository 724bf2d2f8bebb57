// Fused per-problem LTV-MPC step, version 2 (the product kernel).
//
// One CTA per problem (kinematic N = 40: 6 warps, 2 CTAs/SM; N = 20: 4 warps x 5; N = 80: 12 warps with
// the operator tile split between registers and shared memory; dynamic: 8 warps).  The nV x nV dual
// active-set operator M = [K1 | J2] lives in REGISTERS, tiled so that
//     warp w  owns rows    w*RPW .. w*RPW+RPW-1           (RPW = ceil(nV/NW) = 14 for nV = 81, 6 warps)
//     lane l  owns columns l, l+32, l+64                   (CS  = ceil(nV/32) = 3)
// i.e. every thread holds an RPW x CS tile (42 doubles).  With that layout
//   * y = M'n     : 42 FMAs per thread + ONE cross-warp sum through shared memory,
//   * z = J2 y2   : 42 FMAs per thread + an in-warp reduce-scatter (shuffles only),
//   * rank-1 update of M : one FMA per element, no communication,
// and every warp derives step lengths / add-or-drop decisions redundantly from the same
// data, so an add iteration needs 3 block barriers and no serial "warp 0 decides" section.
// Setup without a dense factorisation: a Gramian and a Riccati recursion over the horizon (one
// warp, horizon_recursions*) give every entry of the condensed Hessian as one short dot product
// and J (J'HJ = I) as the closed-loop response to unit innovations, its rows produced by adjoint
// threads that follow the recursion stage by stage.
//
// Pipeline per CTA (reference function each stage replaces):
//   load        x0, x_ref, x_lin, u_lin                      (arguments of ltvmpc_*_curvilinear.m:1)
//   linearise   A_k, B_k, d_k per step                        (rk2_kinematic_curvilinear.m:25-50 ...)
//   discretise  A_k*dt+I, B*dt, d*dt, free response           (sequential_integration.m:16-18,21-26,38-47)
//   condense    packed B_bar rows, H, g, row bounds           (sequential_integration.m:28-36,
//                                                              *_state_constraints.m, generate_qp.m:23-33)
//   solve       Goldfarb-Idnani dual active set, operator form (qpOASES call, ltvmpc_*_curvilinear.m:52)
//   output      u_opt, slack_opt, x_opt, fval, exitflag       (ltvmpc_*_curvilinear.m:57-60)
#pragma once
#include "batch.cuh"   // BatchArgs, Dims
#include "gi_core.cuh"
#include "gi_core_rl.cuh"
#include <type_traits>

// w-space row search of the kinematic model: 2 = lanes of a step pair split between its two rows (default),
// 1 = every lane strides through both rows (selects in the loop).  Measured per horizon, see DESIGN.md.
#ifndef FSAE_SEARCH
#define FSAE_SEARCH 2
#endif

namespace fsae {

template <class Model, int N, int NW_ = 8, int KB_ = 1, int CSR_ = -1>
using CfgV2 = GiCfg<Dims<Model, N>::nV, NW_, KB_, CSR_>;

// LONG (CSR_ >= 0, long horizons): part of the operator tile lives in shared memory (Msm, which
// reuses the space of arrays that are dead once the tiles are filled), only the B_bar rows the
// constraints touch stay in shared memory, and the other B_bar rows, the packed H and the J
// staging live in a per-problem global slab that stays L2-resident.
template <class Model, int N, int NW_ = 8, int KB_ = 1, int CSR_ = -1>
struct SmemV2 {
    using D = Dims<Model, N>;
    using C = Cons<Model>;
    using G = CfgV2<Model, N, NW_, KB_, CSR_>;
    static constexpr bool LONG = G::CSS > 0;
    // BFG: even the B_bar rows the constraints touch do not fit shared memory (dynamic model at horizon 80: 4 x 52 KB);
    // they are read from the problem's L2-resident slab like the other rows
    static constexpr bool BFG = LONG && (size_t)C::NCR * D::NPK * sizeof(double) > 65536;
    static constexpr int NBF = BFG ? 0 : (LONG ? C::NCR : C::NREAL);       // B_bar rows in shared memory
    __host__ __device__ static constexpr int bfc(int c) { return LONG ? c : C::cons_real(c); }   // row of constraint state c
    static constexpr size_t SLAB = LONG ? (size_t)C::NREAL * D::NPK + (size_t)D::nV * D::nV : 0;  // doubles per problem (B_bar rows + full H)
    static_assert(!BFG || (SLAB % 2 == 0 && D::NPK % 2 == 0), "paired (double2) loads from the slab need 16-byte aligned rows");
    alignas(16) double Bf[NBF > 0 ? NBF * D::NPK : 2];   // packed B_bar rows (kept to the end)
    const double* bfg;                          // BFG: this problem's B_bar rows in the slab
    // packed B_bar row of constraint state c
    __device__ __forceinline__ const double* crow(int c) const { return BFG ? bfg + C::cons_real(c) * D::NPK : Bf + bfc(c) * D::NPK; }
    // RL: the row-lane core (gi_core_rl.cuh) for models whose normals are sparse in integrator coordinates
    static constexpr bool RL = C::USE_RL && C::WSPACE && !LONG && KB_ == 1;
    using GR = RlCfg<D::nV, NW_>;
    using Gi_t = std::conditional_t<RL, RlSm<GR, D::NSLOT, D::nU, (D::NS > 0 ? D::nU : -1)>, GiSm<G, D::NSLOT, LONG, C::WSPACE ? D::nU : 0>>;
    Gi_t gi;                                    // x, g, packed H, working set, core scratch
    double B1[D::NX * D::NU];
    double xf[N * D::NX];
    union {
        struct {                                // dead once the operator tiles are filled
            double Ad[N * C::NREAL * D::NX];
            alignas(16) double xl[N * D::NX];  // TMA bulk-copy destinations (16-byte aligned, sizes % 16 == 0)
            alignas(16) double xr[N * D::NX];
            alignas(16) double ul[N * D::NU];
            double dd[N * D::NX];
            double Kc[N * D::NU * D::NX];      // Riccati feedback gains K_s (adjoint rows -> J)
            union {                             // recursion scratch: W, P (double-buffered), W A, P A; B'P, S, K
                double wgram[6 * D::NX * D::NX + 3 * D::NU * D::NX];
                alignas(16) double wgram2[(6 * D::NX + 3 * D::NU) * ((D::NX + 1) & ~1)];   // rows padded to an even length
            };
        };
        alignas(16) double Msm[GiTile<G>::SM_DOUBLES > 0 ? GiTile<G>::SM_DOUBLES : 2];   // shared part of the operator tiles
    };
    double pc[N * C::NPC];
    double g0[N * C::NG0];
    double cg[C::NCG];
    double rlo[D::NROWS], rup[D::NROWS];
    double rn2[D::NROWS];              // squared norm of each row's normal (linear-dependence test)
    static constexpr bool ROWNORMS = true;             // closed form (below): cheap for every model
    static constexpr int NGR = C::NCR * (C::NCR + 1) / 2;
    double gram[ROWNORMS ? N * NGR : 1];            // per step: Gram matrix of the constraint B_bar rows
    double csum[ROWNORMS ? N * C::NCR * D::NU : 1]; // per step: their sums over each control's columns
    double Gs[N * D::NU * D::NX];      // G_s = B' W_{s+1}: cost Gramian seen from the controls of step s
    double Wi[N * D::NU * D::NU];      // (Lambda_s)^(-T/2) / sqrt(2): the diagonal blocks of J
    double Lam3[N * 3];                // Lambda_s (a, b, d)
    static constexpr bool TWOPHASE = C::NR > 8;        // many rows per step: evaluate xs = B_bar_c u first, then the rows
    double xs[TWOPHASE ? C::NXS * N : 1];              // constraint-state perturbations per step (two-phase search)
    double scal[8];                    // 0 cost const
    alignas(8) fsae_params prm;        // this problem's parameter set (copied once: no global loads in the loops)
    alignas(8) unsigned long long mbar; // mbarrier of the input staging
    int prog;                          // last horizon stage the backward recursion has published (N: none yet)
};

// ---- TMA (bulk async copy) staging of one problem's contiguous input records ----------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
    unsigned ok = 0;
    while (!ok) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    }
}

// Problem policy of the core for the LTV-MPC QPs: slots [0, nV) are the variable bounds
// (ltvmpc_*_curvilinear.m:28-29), slots nV + r*N + k the constraint row r at horizon step k
// (cons.cuh).  Nothing dense is ever formed.
template <class Model, int N, int NW_, int KB_, int CSR_>
struct MpcProb {
    using D = Dims<Model, N>;
    using C = Cons<Model>;
    using G = CfgV2<Model, N, NW_, KB_, CSR_>;
    using S_t = SmemV2<Model, N, NW_, KB_, CSR_>;
    S_t& S;
    const fsae_params& P;
    double dt;
    // Candidate reuse between full searches (gi_core.cuh, P1), measured on B200: the dynamic model, whose search
    // evaluates 764 slots, gains 11 % (277 k -> 307 k QP/s at 4 % more iterations); the kinematic model LOSES 16 %
    // (46.7 -> 59.0 iterations per QP: most-violated pivoting matters more than the cheap search is worth), so it
    // stays off there.  -DFSAE_REUSE=0/1 overrides for A/B builds.
#ifdef FSAE_REUSE
    static constexpr bool REUSE = FSAE_REUSE != 0;
#else
    static constexpr bool REUSE = S_t::TWOPHASE;
#endif

    // Exact value of ONE slot side (code = slot * 2 + upper) at the current x, computed by the whole warp:
    // negative = violated by that much.  Same arithmetic as search() for that slot (the lanes split the packed
    // B_bar row by column pairs).
    __device__ __forceinline__ double eval_code(int code) const {
        constexpr int NU = D::NU, nU = D::nU, nV = D::nV;
        const int lane = threadIdx.x & 31;
        const double* x = S.gi.x;
        const int slot = code >> 1;
        const bool upper = code & 1;
        if (slot < nV) {
            const double xv = x[slot];
            const double lb = (slot < nU) ? P.u_lb[slot % NU] : 0.0;
            const double ub = (slot < nU) ? P.u_ub[slot % NU] : INFINITY;
            return upper ? ub - xv : xv - lb;
        }
        const int rr = slot - nV, r = rr / N, k = rr - r * N;
        double acc[C::NXS];
#pragma unroll
        for (int c = 0; c < C::NXS; ++c) acc[c] = 0.0;
        const int len = NU * (k + 1);
        for (int j = 2 * lane; j < len; j += 64) {
            const double2 xx = *reinterpret_cast<const double2*>(&x[j]);
#pragma unroll
            for (int c = 0; c < C::NCR; ++c) {
                const double2 bb = *reinterpret_cast<const double2*>(&S.crow(c)[D::pk(k, j)]);
                acc[c] = fma(bb.x, xx.x, fma(bb.y, xx.y, acc[c]));
            }
#pragma unroll
            for (int ci = 0; ci < C::NINT; ++ci) acc[C::NCR + ci] += (C::int_ucol(ci) == 0) ? xx.x : xx.y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int c = 0; c < C::NXS; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
        }
#pragma unroll
        for (int ci = 0; ci < C::NINT; ++ci) acc[C::NCR + ci] *= dt;
        const double rv = C::row_value(r, acc, S.pc + k * C::NPC, S.cg, x[NU * k]);
        const int sl = C::row_slack(r);
        const double sv = sl >= 0 ? x[nU + sl] : 0.0;
        return upper ? (S.rup[rr] - rv + sv) : (rv + sv - S.rlo[rr]);
    }

    // Threads [0, 4N): 4 lanes per horizon step k compute the constraint-state perturbations
    // xs[., k] = (B_bar_c x)[., k] (packed rows, plus the exact prefix sums of the integrator
    // states) and split that step's rows among themselves; threads [4N, 4N+nV): the bounds.
    __device__ __forceinline__ void search(double& best, int& best_i) const {
        constexpr int NU = D::NU, nU = D::nU, nV = D::nV, NT = G::NT;
        constexpr int ROWT = ((4 * N + 31) / 32) * 32;      // row threads, rounded up to whole warps
        static_assert(NU == 2, "paired (double2) row loads assume two controls per step");
        const int tid = threadIdx.x;
        const double* x = S.gi.x;
        if constexpr (S_t::TWOPHASE) {
            // Many rows per step (dynamic model: 17): (a) the NCR x N dot products xs[c][k] = B_bar_c[k] . u
            // spread over ALL threads -- two lanes for the long ones (k >= K2), one for the short --
            // then (b) the NROWS + nV slots dealt round-robin.  One extra barrier, balanced work.
            constexpr int NCR = C::NCR;
            constexpr int K2 = (2 * N - NT / NCR) > 0 ? (2 * N - NT / NCR) : 0;
            static_assert(K2 <= N && (2 * (N - K2) + K2) * NCR <= NT, "two-phase search thread map");
            constexpr int NP2 = (N - K2) * NCR;             // dot products with two lanes
            static_assert((2 * NP2) % 32 == 0, "the two-lane dot products fill whole warps");
            int c = -1, k = 0, half = 0, lanes = 1;
            if (tid < 2 * NP2) { const int p = tid >> 1; half = tid & 1; lanes = 2; c = p / (N - K2); k = K2 + p - c * (N - K2); }
            else if (K2 > 0 && tid - 2 * NP2 < K2 * NCR) { const int q = tid - 2 * NP2; c = q / K2; k = q - c * K2; }
            double acc = 0.0, ai[C::NINT > 0 ? C::NINT : 1];
#pragma unroll
            for (int ci = 0; ci < C::NINT; ++ci) ai[ci] = 0.0;
            if (c >= 0) {
                const int np = k + 1;                         // double2 pairs of this row
                const int h = (lanes == 2) ? (np + 1) >> 1 : np;
                const int p0 = half * h, p1 = (p0 + h < np) ? p0 + h : np;
                const double* brow = S.crow(c) + D::pk(k, 0);
                double a1 = 0.0;
                for (int pp = p0; pp < p1; ++pp) {
                    const double2 xx = *reinterpret_cast<const double2*>(&x[2 * pp]);
                    const double2 bb = *reinterpret_cast<const double2*>(&brow[2 * pp]);
                    acc = fma(bb.x, xx.x, acc);
                    a1 = fma(bb.y, xx.y, a1);
                    if (c == 0) {
#pragma unroll
                        for (int ci = 0; ci < C::NINT; ++ci) ai[ci] += (C::int_ucol(ci) == 0) ? xx.x : xx.y;
                    }
                }
                acc += a1;
            }
            if (tid < 2 * NP2) {                             // whole warps (2 * NP2 is a multiple of 32 for the supported sizes)
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
#pragma unroll
                for (int ci = 0; ci < C::NINT; ++ci) ai[ci] += __shfl_xor_sync(0xffffffffu, ai[ci], 1);
            }
            if (c >= 0 && half == 0) {
                S.xs[c * N + k] = acc;
                if (c == 0) {
#pragma unroll
                    for (int ci = 0; ci < C::NINT; ++ci) S.xs[(NCR + ci) * N + k] = ai[ci] * dt;
                }
            }
            __syncthreads();
            for (int sl_ = tid; sl_ < D::NROWS + nV; sl_ += NT) {
                if (sl_ < D::NROWS) {
                    const int r = sl_ / N, kk = sl_ - r * N, slot = nV + sl_;
                    if (S.gi.status[slot] != 0) continue;
                    double xsk[C::NXS];
#pragma unroll
                    for (int cc = 0; cc < C::NXS; ++cc) xsk[cc] = S.xs[cc * N + kk];
                    const double rv = C::row_value(r, xsk, S.pc + kk * C::NPC, S.cg, x[NU * kk]);
                    const int sl = C::row_slack(r);
                    const double sv = sl >= 0 ? x[nU + sl] : 0.0;
                    const double vlo = rv + sv - S.rlo[sl_];
                    const double vup = S.rup[sl_] - rv + sv;
                    if (vlo < best) { best = vlo; best_i = slot * 2; }
                    if (vup < best) { best = vup; best_i = slot * 2 + 1; }
                } else {
                    const int slot = sl_ - D::NROWS;
                    if (S.gi.status[slot] != 0) continue;
                    const double xv = x[slot];
                    const double lb = (slot < nU) ? P.u_lb[slot % NU] : 0.0;
                    const double ub = (slot < nU) ? P.u_ub[slot % NU] : INFINITY;
                    const double vlo = xv - lb, vup = ub - xv;
                    if (vlo < best) { best = vlo; best_i = slot * 2; }
                    if (vup < best) { best = vup; best_i = slot * 2 + 1; }
                }
            }
            return;
        }
        if constexpr (C::WSPACE) {
            // Integrator coordinates: the v / delta perturbations of step k ARE x[2k], x[2k+1]; only the rows of the
            // real states need a dot product (packed rows transformed to w-space at setup).  Eight lanes per PAIR of
            // steps (k, N-1-k): the two packed rows of a pair hold N+1 column pairs together, so every group does
            // the same work; then each of the eight lanes evaluates one of the pair's 2 x NR rows.
            static_assert(N % 2 == 0 && C::NR == 4, "step pairs, one row per lane");
            // The eight lanes of a group are split between the two rows in proportion to their lengths (nA lanes for
            // row ka, the others for row kb), each lane striding through ITS row: one accumulator, no selects in the
            // loop, at most NITER column pairs per lane.
            constexpr int NITER = (N + 1 + 6) / 7;
            const double idt = S.gi.idt;
            for (int rt = tid; rt < ROWT; rt += NT) {        // warp-uniform trip count
                const int gq = rt >> 3, part = rt & 7;
                const bool valid = gq < N / 2;
                const int ka = valid ? gq : 0, kb = N - 1 - ka;
#if FSAE_SEARCH == 1
                double accA[C::NCR], accB[C::NCR];
#pragma unroll
                for (int c = 0; c < C::NCR; ++c) { accA[c] = 0.0; accB[c] = 0.0; }
#pragma unroll
                for (int it = 0; it < (N + 1 + 7) / 8; ++it) {       // lane `part` takes every eighth column pair of the two rows
                    const int p = part + 8 * it;
                    if (p < N + 1) {
                        const bool inA = p <= ka;
                        const int kk = inA ? ka : kb, pp = inA ? p : p - ka - 1;
                        const double2 xx = *reinterpret_cast<const double2*>(&x[2 * pp]);
#pragma unroll
                        for (int c = 0; c < C::NCR; ++c) {
                            const double2 bb = *reinterpret_cast<const double2*>(&S.crow(c)[D::pk(kk, 2 * pp)]);
                            const double t = fma(bb.x, xx.x, bb.y * xx.y);
                            accA[c] += inA ? t : 0.0;
                            accB[c] += inA ? 0.0 : t;
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < C::NCR; ++c) {
#pragma unroll
                    for (int o = 1; o < 8; o <<= 1) {
                        accA[c] += __shfl_xor_sync(0xffffffffu, accA[c], o);
                        accB[c] += __shfl_xor_sync(0xffffffffu, accB[c], o);
                    }
                }
#else
                const int nA = (ka + NITER) / NITER;          // ceil((ka + 1) / NITER) lanes for row ka
                const bool isA = part < nA;
                const int kk = isA ? ka : kb;
                const int first = isA ? part : part - nA, stride = isA ? nA : 8 - nA, len = kk + 1;
                double acc[C::NCR];
#pragma unroll
                for (int c = 0; c < C::NCR; ++c) acc[c] = 0.0;
                const double2* x2 = reinterpret_cast<const double2*>(x);
#pragma unroll
                for (int it = 0; it < NITER; ++it) {
                    const int pp = first + it * stride;
                    if (pp < len) {
                        const double2 xx = x2[pp];
#pragma unroll
                        for (int c = 0; c < C::NCR; ++c) {
                            const double2 bb = reinterpret_cast<const double2*>(S.crow(c) + D::pk(kk, 0))[pp];
                            acc[c] = fma(bb.y, xx.y, fma(bb.x, xx.x, acc[c]));
                        }
                    }
                }
                double accA[C::NCR], accB[C::NCR];
#pragma unroll
                for (int c = 0; c < C::NCR; ++c) {
                    accA[c] = isA ? acc[c] : 0.0;
                    accB[c] = isA ? 0.0 : acc[c];
#pragma unroll
                    for (int o = 1; o < 8; o <<= 1) {
                        accA[c] += __shfl_xor_sync(0xffffffffu, accA[c], o);
                        accB[c] += __shfl_xor_sync(0xffffffffu, accB[c], o);
                    }
                }
#endif
                if (valid) {
                    const int k = part < 4 ? ka : kb, r = part & 3;
                    const int rr = r * N + k, slot = nV + rr;
                    if (S.gi.status[slot] == 0) {
                        double av[C::NXS];
#pragma unroll
                        for (int c = 0; c < C::NCR; ++c) av[c] = part < 4 ? accA[c] : accB[c];
#pragma unroll
                        for (int ci = 0; ci < C::NINT; ++ci) av[C::NCR + ci] = x[NU * k + C::int_ucol(ci)];
                        const double rv = C::row_value(r, av, S.pc + k * C::NPC, S.cg, 0.0);
                        const int sl = C::row_slack(r);
                        const double sv = sl >= 0 ? x[nU + sl] : 0.0;
                        const double vlo = rv + sv - S.rlo[rr];
                        const double vup = S.rup[rr] - rv + sv;
                        if (vlo < best) { best = vlo; best_i = slot * 2; }
                        if (vup < best) { best = vup; best_i = slot * 2 + 1; }
                    }
                }
            }
            // variable bounds: u_i = (w_i - w_{i-2}) / dt for the controls, the slacks as they are
            for (int slot = (tid + NT - ROWT % NT) % NT; slot < nV; slot += NT) {
                if (S.gi.status[slot] == 0) {
                    const double xv = (slot < nU) ? (x[slot] - (slot >= NU ? x[slot - NU] : 0.0)) * idt : x[slot];
                    const double lb = (slot < nU) ? P.u_lb[slot % NU] : 0.0;
                    const double ub = (slot < nU) ? P.u_ub[slot % NU] : INFINITY;
                    const double vlo = xv - lb, vup = ub - xv;
                    if (vlo < best) { best = vlo; best_i = slot * 2; }
                    if (vup < best) { best = vup; best_i = slot * 2 + 1; }
                }
            }
            return;
        }
        for (int rt = tid; rt < ROWT; rt += NT) {            // warp-uniform trip count
            const int k = rt >> 2, part = rt & 3;
            const bool valid = k < N;
            double acc[C::NXS];
#pragma unroll
            for (int c = 0; c < C::NXS; ++c) acc[c] = 0.0;
            if (valid) {
                const int len = NU * (k + 1);
                // the four lanes of a step take the column pairs round-robin: consecutive lanes read consecutive
                // 16-byte pairs of the packed row and of x (a chunk per lane put the lanes a chunk apart: bank conflicts)
                double accb[C::NCR];                     // second partial sum per row: two 8-cycle chains instead of one of 16
#pragma unroll
                for (int c = 0; c < C::NCR; ++c) accb[c] = 0.0;
#pragma unroll 2
                for (int j = 2 * part; j < len; j += 8) {
                    const double2 xx = *reinterpret_cast<const double2*>(&x[j]);
#pragma unroll
                    for (int c = 0; c < C::NCR; ++c) {
                        const double2 bb = *reinterpret_cast<const double2*>(&S.crow(c)[D::pk(k, j)]);
                        acc[c] = fma(bb.x, xx.x, acc[c]);
                        accb[c] = fma(bb.y, xx.y, accb[c]);
                    }
#pragma unroll
                    for (int ci = 0; ci < C::NINT; ++ci) acc[C::NCR + ci] += (C::int_ucol(ci) == 0) ? xx.x : xx.y;
                }
#pragma unroll
                for (int c = 0; c < C::NCR; ++c) acc[c] += accb[c];
            }
#pragma unroll
            for (int c = 0; c < C::NXS; ++c) {
                acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 1);
                acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 2);
            }
            if (valid) {
#pragma unroll
                for (int ci = 0; ci < C::NINT; ++ci) acc[C::NCR + ci] *= dt;
                const double ua = x[NU * k];
                for (int r = part; r < C::NR; r += 4) {
                    const int rr = r * N + k, slot = nV + rr;
                    if (S.gi.status[slot] != 0) continue;
                    const double rv = C::row_value(r, acc, S.pc + k * C::NPC, S.cg, ua);
                    const int sl = C::row_slack(r);
                    const double sv = sl >= 0 ? x[nU + sl] : 0.0;
                    const double vlo = rv + sv - S.rlo[rr];
                    const double vup = S.rup[rr] - rv + sv;
                    if (vlo < best) { best = vlo; best_i = slot * 2; }
                    if (vup < best) { best = vup; best_i = slot * 2 + 1; }
                }
            }
        }
        // variable bounds: dealt round-robin starting at the first thread after the row threads
        // (with 256 threads: one slot per thread of the three bound warps, as before)
        for (int slot = (tid + NT - ROWT % NT) % NT; slot < nV; slot += NT) {
            if (S.gi.status[slot] == 0) {
                const double xv = x[slot];
                const double lb = (slot < nU) ? P.u_lb[slot % NU] : 0.0;
                const double ub = (slot < nU) ? P.u_ub[slot % NU] : INFINITY;
                const double vlo = xv - lb, vup = ub - xv;
                if (vlo < best) { best = vlo; best_i = slot * 2; }
                if (vup < best) { best = vup; best_i = slot * 2 + 1; }
            }
        }
    }

    // Normal of (slot, side) in  n'x >= b  form.  Everything that depends only on the slot is
    // computed once, warp-uniformly (no divergence); the per-entry part is one shared-memory
    // load and selects.
    // Normal with at most three entries (ascending), or 0 = dense.  u coordinates: the variable bounds are unit
    // normals.  Integrator coordinates (C::WSPACE): a control bound is (w_i - w_{i-2}) / dt, rows that touch only
    // integrator states and a slack have their two or three entries (cons.cuh, sparse_row).
    __device__ __forceinline__ SpN sparse_normal(int pslot, int pside) const {
        constexpr int NU = D::NU, nU = D::nU, nV = D::nV;
        const double sg = pside < 0 ? 1.0 : -1.0;
        if (pslot < nV) {
            if (C::WSPACE && pslot < nU) {
                const double c = sg * S.gi.idt;
                if (pslot < NU) return SpN{1, pslot, 0, 0, c, 0.0, 0.0};
                return SpN{2, pslot - NU, pslot, 0, -c, c, 0.0};
            }
            return SpN{1, pslot, 0, 0, sg, 0.0, 0.0};
        }
        if (!C::WSPACE) return spn_none();
        const int rr = pslot - nV, r = rr / N, k = rr - r * N;
        return C::sparse_row(r, k, S.pc + k * C::NPC, sg, nU);
    }

    struct Prep {
        int pslot, k;            // k = horizon step of a row slot, -1 for a variable bound
        SpN sn;                  // integrator coordinates: sparse normals (dense fallback of the block variants)
        double sg;
        double creal[C::NCR];    // coefficients of the packed B_bar rows
        double cctl[D::NU];      // dt * (integrator-state coefficients) per control column
        double cu[D::NU];        // direct control coefficients at step k
        int slack;               // absolute index of the slack entry, -1 if none
    };
    __device__ __forceinline__ Prep normal_prepare(int pslot, int pside) const {
        constexpr int NU = D::NU, nU = D::nU, nV = D::nV;
        Prep p;
        p.pslot = pslot;
        p.sg = pside < 0 ? 1.0 : -1.0;
        p.k = -1;
        p.slack = -1;
        p.sn = C::WSPACE ? sparse_normal(pslot, pside) : spn_none();
#pragma unroll
        for (int c = 0; c < C::NCR; ++c) p.creal[c] = 0.0;
#pragma unroll
        for (int c = 0; c < NU; ++c) { p.cctl[c] = 0.0; p.cu[c] = 0.0; }
        if (pslot >= nV) {
            const int rr = pslot - nV, r = rr / N, k = rr - r * N;
            const double* pc = S.pc + k * C::NPC;
            p.k = k;
            const int sl = C::row_slack(r);
            p.slack = sl >= 0 ? nU + sl : -1;
#pragma unroll
            for (int c = 0; c < C::NCR; ++c) p.creal[c] = p.sg * C::row_coef(r, c, pc, S.cg);
#pragma unroll
            for (int c = 0; c < C::NINT; ++c) p.cctl[C::int_ucol(c)] += p.sg * (C::WSPACE ? 1.0 : dt) * C::row_coef(r, C::NCR + c, pc, S.cg);
#pragma unroll
            for (int c = 0; c < NU; ++c) p.cu[c] = p.sg * C::row_ucoef(r, c, pc, S.cg);
        }
        return p;
    }
    __device__ __forceinline__ double normal_entry(const Prep& p, int i) const {
        constexpr int NU = D::NU, nU = D::nU;
        if constexpr (C::WSPACE) {
            if (p.sn.cnt > 0) {                                         // uniform branch
                double v = 0.0;
#pragma unroll
                for (int e = 0; e < 3; ++e) v += (e < p.sn.cnt && i == p.sn.idx(e)) ? p.sn.cf(e) : 0.0;
                return v;
            }
            // dense row in integrator coordinates: the transformed packed rows, the integrator states' own entries
            const int step = i / NU, uc = i - step * NU;
            const bool in = (i < nU) & (step <= p.k);
            double acc = (step == p.k) ? ((uc == 0) ? p.cctl[0] : p.cctl[NU - 1]) : 0.0;   // (no direct control terms in these models)
            const int idx = in ? D::pk(p.k, i) : 0;
#pragma unroll
            for (int c = 0; c < C::NCR; ++c) acc = fma(p.creal[c], S.crow(c)[idx], acc);
            return in ? acc : ((i == p.slack) ? 1.0 : 0.0);
        }
        if (p.k < 0) return (i == p.pslot) ? p.sg : 0.0;               // uniform branch
        const int step = i / NU, uc = i - step * NU;
        const bool in = (i < nU) & (step <= p.k);
        double acc = (uc == 0) ? p.cctl[0] : p.cctl[NU - 1];
        acc += (step == p.k) ? ((uc == 0) ? p.cu[0] : p.cu[NU - 1]) : 0.0;
        const int idx = in ? D::pk(p.k, i) : 0;
#pragma unroll
        for (int c = 0; c < C::NCR; ++c) acc = fma(p.creal[c], S.crow(c)[idx], acc);
        return in ? acc : ((i == p.slack) ? 1.0 : 0.0);
    }

    __device__ __forceinline__ double norm2(int pslot) const {
        return (pslot < D::nV) ? 1.0 : S.rn2[pslot - D::nV];
    }
    __device__ __forceinline__ bool is_unit(int pslot) const { return pslot < D::nV && !(C::WSPACE && pslot < D::nU); }
};

// The two backward recursions over the horizon (s = N-1 .. 0), run by ONE warp each (or both by the same
// warp when the CTA has only four):
//   DO_W  Gramian  W_s = Q + A_s' W_{s+1} A_s                       -> G_s = B' W_{s+1}   (entries of H)
//   DO_P  Riccati  P_s = Q + A_s' P_{s+1} A_s - S' Lambda^-1 S,  Lambda = R + B' P_{s+1} B,
//                  S = B' P_{s+1} A_s,  K_s = -Lambda^-1 S           -> J (adjoint rows), stage published
// The Riccati identity  u'(B_bar' Qbar B_bar + Rbar) u = sum_s |Lambda_s^(1/2) (u_s - K_s x_s)|^2
// makes J = T^-1 / sqrt 2 (J'HJ = I) the closed-loop response to unit "innovations":
// u_s = K_s x_s + Lambda_s^(-T/2) v_s.  No dense factorisation of H is needed.
// Runtime horizon Na <= N (the template's capacity): steps k >= Na are padding -- zero state cost, control cost R
// (their controls stay exactly 0 and decouple), rows disabled.  State weight of predicted step k (0-based):
__device__ __forceinline__ double stage_weight(const fsae_params& P, int k, int r, int Na) {
    return k < Na - 1 ? P.Q[r] : (k == Na - 1 ? P.Q_terminal[r] : 0.0);
}

template <bool DO_W, bool DO_P, class Model, int N, class S_t>
__device__ __forceinline__ void horizon_recursions(S_t& S, const fsae_params& P, int Na) {
    using D = Dims<Model, N>;
    using C = Cons<Model>;
    constexpr int NX = D::NX, NU = D::NU;
    const int lane = threadIdx.x & 31;
    {
        static_assert(NU == 2, "2 x 2 Lambda blocks are inverted in closed form");
        constexpr int NN = NX * NX, NB = NU * NX, NR_ = C::NREAL;
        double* Wb = S.wgram;               // [2][NN]
        double* Pb = S.wgram + 2 * NN;      // [2][NN]
        double* TW = S.wgram + 4 * NN;      // W A
        double* TP = S.wgram + 5 * NN;      // P A
        double* BP = S.wgram + 6 * NN;      // B'P            [NU][NX]
        double* Sm = BP + NB;               // S = B'P A      [NU][NX]
        double* Kx = Sm + NB;               // K              [NU][NX]
        for (int e = lane; e < NN; e += 32) {
            const double v = (e / NX == e % NX) ? stage_weight(P, N - 1, e / NX, Na) : 0.0;
            if (DO_W) Wb[e] = v;
            if (DO_P) Pb[e] = v;
        }
        __syncwarp();
        int cur = 0;
        RSTAGE_DECL;
        // Three warp-synchronous phases per stage; every lane runs the same instruction stream on its own
        // entry (indices clamped, stores predicated), the dot products of a phase are independent chains.
        for (int st = N - 1; st >= 0; --st) {
            const double* W = Wb + cur * NN;
            const double* Pm = Pb + cur * NN;
            const double* As = S.Ad + st * NR_ * NX;          // rows of the real states; integrator rows are unit rows
            // phase 1: W A, P A (entry (i, j));  B'P, G_s = B'W (entry (c = i, j), i < NU)
            for (int e0 = 0; e0 < NN; e0 += 32) {
                const int e = (e0 + lane < NN) ? e0 + lane : 0;
                const int i = e / NX, j = e - i * NX, c = i < NU ? i : NU - 1;
                double tw = (DO_W && j >= NR_) ? W[i * NX + j] : 0.0, tp = (DO_P && j >= NR_) ? Pm[i * NX + j] : 0.0, bp = 0.0, wb = 0.0;
#pragma unroll
                for (int l = 0; l < NX; ++l) {
                    const double bl = S.B1[l * NU + c];
                    if (DO_P) bp = fma(bl, Pm[l * NX + j], bp);
                    if (DO_W) wb = fma(bl, W[l * NX + j], wb);
                    if (l < NR_) {
                        const double al = As[l * NX + j];
                        if (DO_W) tw = fma(W[i * NX + l], al, tw);
                        if (DO_P) tp = fma(Pm[i * NX + l], al, tp);
                    }
                }
                if (e0 + lane < NN) { if (DO_W) TW[e] = tw; if (DO_P) TP[e] = tp; }
                if (e0 + lane < NB) { if (DO_P) BP[e] = bp; if (DO_W) S.Gs[(st * NU + c) * NX + j] = wb; }
            }
            __syncwarp();
            RSTAGE(0);
            // phase 2 (lane = (c, j), c < NU): S = (B'P) A column j, Lambda = R + (B'P) B, K = -Lambda^-1 S
            if (DO_P) {
                const int e = lane < NB ? lane : 0;
                const int c = e / NX, j = e - c * NX;
                double s0 = (j >= NR_) ? BP[j] : 0.0, s1 = (j >= NR_) ? BP[NX + j] : 0.0;
                double la = P.R[0], lb = 0.0, ld = P.R[1];
#pragma unroll
                for (int l = 0; l < NX; ++l) {
                    const double b0 = BP[l], b1 = BP[NX + l];
                    la = fma(b0, S.B1[l * NU + 0], la);
                    lb = fma(b0, S.B1[l * NU + 1], lb);
                    ld = fma(b1, S.B1[l * NU + 1], ld);
                    if (l < NR_) {
                        const double al = As[l * NX + j];
                        s0 = fma(b0, al, s0);
                        s1 = fma(b1, al, s1);
                    }
                }
                const double rdet = __drcp_rn(fma(la, ld, -lb * lb));
                const double kv = (c == 0) ? (lb * s1 - ld * s0) * rdet : (lb * s0 - la * s1) * rdet;
                if (lane < NB) {
                    Sm[e] = (c == 0) ? s0 : s1;
                    Kx[e] = kv;
                    S.Kc[(st * NU + c) * NX + j] = kv;
                }
                if (lane < 3) S.Lam3[st * 3 + lane] = (lane == 0) ? la : (lane == 1 ? lb : ld);
            }
            if (DO_P) {
                __syncwarp();
                if (lane == 0) {                   // K_st is in shared memory: publish the stage
                    __threadfence_block();
                    *(volatile int*)&S.prog = st;
                }
            }
            RSTAGE(1);
            if (st == 0) break;
            // phase 3: W' = Q + A'(W A),  P' = Q + A'(P A) + S'K
            double* Wn = Wb + (cur ^ 1) * NN;
            double* Pn = Pb + (cur ^ 1) * NN;
            for (int e0 = 0; e0 < NN; e0 += 32) {
                const int e = (e0 + lane < NN) ? e0 + lane : 0;
                const int i = e / NX, j = e - i * NX;
                const double qd = (i == j) ? stage_weight(P, st - 1, i, Na) : 0.0;
                double aw = qd + ((DO_W && i >= NR_) ? TW[i * NX + j] : 0.0);
                double ap = qd + ((DO_P && i >= NR_) ? TP[i * NX + j] : 0.0);
                double sk = 0.0;
#pragma unroll
                for (int l = 0; l < NR_; ++l) {
                    const double al = As[l * NX + i];
                    if (DO_W) aw = fma(al, TW[l * NX + j], aw);
                    if (DO_P) ap = fma(al, TP[l * NX + j], ap);
                }
                if (DO_P) {
#pragma unroll
                    for (int c = 0; c < NU; ++c) sk = fma(Sm[c * NX + i], Kx[c * NX + j], sk);
                }
                if (e0 + lane < NN) { if (DO_W) Wn[e] = aw; if (DO_P) Pn[e] = ap + sk; }
            }
            __syncwarp();
            RSTAGE(2);
            cur ^= 1;
        }
        // Lambda_s^(-T/2) / sqrt 2 for every stage (off the recursion's critical path).  Lambda = G G',
        // G = [l11 0; l21 l22]:  G^-T = [1/l11  -l21/(l11 l22); 0  1/l22]
        for (int st = lane; DO_P && st < N; st += 32) {
            const double la = S.Lam3[st * 3], lb = S.Lam3[st * 3 + 1], ld = S.Lam3[st * 3 + 2];
            const double i11 = rsqrt(la), l21 = lb * i11;
            const double i22 = rsqrt(fma(-l21, l21, ld));
            const double r2 = 0.70710678118654752440;
            S.Wi[st * 4 + 0] = r2 * i11;
            S.Wi[st * 4 + 1] = -r2 * l21 * i11 * i22;
            S.Wi[st * 4 + 2] = 0.0;
            S.Wi[st * 4 + 3] = r2 * i22;
        }
    }
}

// The same two recursions for small state dimensions (NX^2 <= 32: one matrix entry per lane), tuned for the
// critical path of the kinematic kernel: rows padded to an even length and read as double2 (W, P are
// symmetric, so columns are read as rows; W A and P A are stored transposed for the same reason), B and the
// lane's column of A_s in registers.  ~40 shared-memory loads per stage instead of ~66.
template <class Model, int N, class S_t>
__device__ __forceinline__ void horizon_recursions_small(S_t& S, const fsae_params& P, int Na) {
    using D = Dims<Model, N>;
    using C = Cons<Model>;
    constexpr int NX = D::NX, NU = D::NU, NR_ = C::NREAL, NXP = (NX + 1) & ~1, NV2 = NXP / 2;
    static_assert(NX * NX <= 32 && NU == 2, "one matrix entry per lane, 2 x 2 Lambda blocks");
    const int lane = threadIdx.x & 31;
    const int e = lane < NX * NX ? lane : 0;
    const int i = e / NX, j = e - i * NX, c = i < NU ? i : NU - 1;
    const bool le = lane < NX * NX, lb_ = lane < NU * NX;
    // scratch (doubles): W [2][NX][NXP], P [2][NX][NXP], (W A)' [NX][NXP], (P A)' [NX][NXP], B'P [NU][NXP], S [NU][NXP], K [NU][NXP]
    double* Wb = S.wgram2;
    double* Pb = Wb + 2 * NX * NXP;
    double* TWt = Pb + 2 * NX * NXP;
    double* TPt = TWt + NX * NXP;
    double* BP = TPt + NX * NXP;
    double* Sm = BP + NU * NXP;
    double* Kx = Sm + NU * NXP;
    auto ldrow = [](const double* M, int r, double (&o)[NXP]) {
        const double2* p2 = reinterpret_cast<const double2*>(M + r * NXP);
#pragma unroll
        for (int h = 0; h < NV2; ++h) { const double2 v = p2[h]; o[2 * h] = v.x; o[2 * h + 1] = v.y; }
    };
    double bc[NX], bb0[NX], bb1[NX];
#pragma unroll
    for (int l = 0; l < NX; ++l) { bc[l] = S.B1[l * NU + c]; bb0[l] = S.B1[l * NU]; bb1[l] = S.B1[l * NU + 1]; }
    const double r0 = P.R[0], r1 = P.R[1];
    for (int t = lane; t < 2 * NX * NXP; t += 32) {
        const int rr = (t % (NX * NXP)) / NXP, cc = t % NXP;
        const double v = (rr == cc) ? stage_weight(P, N - 1, rr, Na) : 0.0;
        Wb[t] = (t < NX * NXP) ? v : 0.0;
        Pb[t] = (t < NX * NXP) ? v : 0.0;
    }
    for (int t = lane; t < 2 * NX * NXP; t += 32) { TWt[t] = 0.0; }      // (W A)', (P A)' incl. padding
    __syncwarp();
    int cur = 0;
    RSTAGE_DECL;
    for (int st = N - 1; st >= 0; --st) {
        const double* W = Wb + cur * NX * NXP;
        const double* Pm = Pb + cur * NX * NXP;
        const double* As = S.Ad + st * NR_ * NX;
        double aj[NR_];
#pragma unroll
        for (int l = 0; l < NR_; ++l) aj[l] = As[l * NX + j];
        // phase 1: (W A)[i][j], (P A)[i][j];  (B'P)[c][j], G_s[c][j] = (B'W)[c][j]
        {
            double wi[NXP], wj[NXP], pi[NXP], pj[NXP];
            ldrow(W, i, wi); ldrow(W, j, wj); ldrow(Pm, i, pi); ldrow(Pm, j, pj);
            double tw = (j >= NR_) ? W[i * NXP + j] : 0.0, tp = (j >= NR_) ? Pm[i * NXP + j] : 0.0, bp = 0.0, wb = 0.0;
#pragma unroll
            for (int l = 0; l < NX; ++l) {
                bp = fma(bc[l], pj[l], bp);
                wb = fma(bc[l], wj[l], wb);
                if (l < NR_) { tw = fma(wi[l], aj[l], tw); tp = fma(pi[l], aj[l], tp); }
            }
            if (le) { TWt[j * NXP + i] = tw; TPt[j * NXP + i] = tp; }
            if (lb_) { BP[c * NXP + j] = bp; S.Gs[(st * NU + c) * NX + j] = wb; }
        }
        __syncwarp();
        RSTAGE(0);
        // phase 2 (lane = (c, j)): S = (B'P) A column j, Lambda = R + (B'P) B, K = -Lambda^-1 S
        {
            double b0[NXP], b1[NXP];
            ldrow(BP, 0, b0); ldrow(BP, 1, b1);
            double s0 = (j >= NR_) ? BP[j] : 0.0, s1 = (j >= NR_) ? BP[NXP + j] : 0.0;
            double la = r0, lb = 0.0, ld = r1;
#pragma unroll
            for (int l = 0; l < NX; ++l) {
                la = fma(b0[l], bb0[l], la);
                lb = fma(b0[l], bb1[l], lb);
                ld = fma(b1[l], bb1[l], ld);
                if (l < NR_) { s0 = fma(b0[l], aj[l], s0); s1 = fma(b1[l], aj[l], s1); }
            }
            const double rdet = __drcp_rn(fma(la, ld, -lb * lb));
            const double kv = (c == 0) ? (lb * s1 - ld * s0) * rdet : (lb * s0 - la * s1) * rdet;
            if (lb_) {
                Sm[c * NXP + j] = (c == 0) ? s0 : s1;
                Kx[c * NXP + j] = kv;
                S.Kc[(st * NU + c) * NX + j] = kv;
            }
            if (lane < 3) S.Lam3[st * 3 + lane] = (lane == 0) ? la : (lane == 1 ? lb : ld);
        }
        __syncwarp();
        if (lane == 0) {                           // K_st and G_st are in shared memory: publish the stage
            __threadfence_block();
            *(volatile int*)&S.prog = st;
        }
        RSTAGE(1);
        if (st == 0) break;
        // phase 3: W'[i][j] = Q + (A'(W A))[i][j],  P'[i][j] = Q + (A'(P A))[i][j] + (S'K)[i][j]
        {
            double tw[NXP], tp[NXP];
            ldrow(TWt, j, tw); ldrow(TPt, j, tp);
            const double qd = (i == j) ? stage_weight(P, st - 1, i, Na) : 0.0;
            double aw = qd + ((i >= NR_) ? TWt[j * NXP + i] : 0.0);
            double ap = qd + ((i >= NR_) ? TPt[j * NXP + i] : 0.0);
#pragma unroll
            for (int l = 0; l < NR_; ++l) {
                const double al = As[l * NX + i];
                aw = fma(al, tw[l], aw);
                ap = fma(al, tp[l], ap);
            }
            double sk = 0.0;
#pragma unroll
            for (int cc = 0; cc < NU; ++cc) sk = fma(Sm[cc * NXP + i], Kx[cc * NXP + j], sk);
            if (le) {
                Wb[(cur ^ 1) * NX * NXP + i * NXP + j] = aw;
                Pb[(cur ^ 1) * NX * NXP + i * NXP + j] = ap + sk;
            }
        }
        __syncwarp();
        RSTAGE(2);
        cur ^= 1;
    }
    for (int st = lane; st < N; st += 32) {
        const double la = S.Lam3[st * 3], lb = S.Lam3[st * 3 + 1], ld = S.Lam3[st * 3 + 2];
        const double i11 = rsqrt(la), l21 = lb * i11;
        const double i22 = rsqrt(fma(-l21, l21, ld));
        const double r2 = 0.70710678118654752440;
        S.Wi[st * 4 + 0] = r2 * i11;
        S.Wi[st * 4 + 1] = -r2 * l21 * i11 * i22;
        S.Wi[st * 4 + 2] = 0.0;
        S.Wi[st * 4 + 3] = r2 * i22;
    }
}

// PAD = false: the problem's horizon IS the capacity N (everything about the horizon folds at compile time);
// PAD = true : runtime horizon a.N < N, the remaining steps are padding.
template <class Model, int N, int MINB, int NW_ = 8, int KB_ = 1, int CSR_ = -1, bool PAD = false>
__global__ void __launch_bounds__(32 * NW_, MINB) ltvmpc_fused_v2_kernel(BatchArgs a) {
    using D = Dims<Model, N>;
    using C = Cons<Model>;
    using G = CfgV2<Model, N, NW_, KB_, CSR_>;
    using S_t = SmemV2<Model, N, NW_, KB_, CSR_>;
    constexpr bool LONG = S_t::LONG;
    constexpr int NX = D::NX, NU = D::NU, NS = D::NS, nU = D::nU, nV = D::nV;
    constexpr int NT = G::NT, NW = G::NW, RPW = G::RPW, CS = G::CS, RP = G::RP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    S_t& S = *reinterpret_cast<S_t*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    if (b >= a.B) return;
    const fsae_params& Pg = a.params[a.param_id ? a.param_id[b] : 0];
    static_assert(sizeof(fsae_params) % 8 == 0, "parameter set is copied as 64-bit words");
    for (int i = tid; i < (int)(sizeof(fsae_params) / 8); i += 32 * NW_)
        reinterpret_cast<unsigned long long*>(&S.prm)[i] = reinterpret_cast<const unsigned long long*>(&Pg)[i];
    const fsae_params& P = S.prm;      // visible after the barrier that ends the load stage
    const DevTrack tr = a.tracks[a.track_id ? a.track_id[b] : 0];
    const double dt = a.dt;
    const int Na = PAD ? a.N : N;           // horizon (1 <= Na <= N): steps >= Na are padding
    const int row0 = warp * RPW;            // first row of this warp
    // all real-state rows of B_bar: shared memory, or (LONG) the problem's global slab [B_bar rows | packed H]
    double* const bf_all = LONG ? a.m_scratch + (size_t)b * S_t::SLAB : S.Bf;
    if (C::WSPACE && tid == 0) S.gi.idt = 1.0 / dt;
    if constexpr (LONG) if (tid == 0) {
        S.gi.hpg = a.m_scratch + (size_t)b * S_t::SLAB + (size_t)C::NREAL * D::NPK;
        S.bfg = bf_all;
    }

    STAGE_DECL;
    // ---------------------------------------------------------------- load (TMA bulk copies)
    // x_lin, u_lin, x_ref of one problem are contiguous records: one elected thread issues
    // three cp.async.bulk copies that complete on an mbarrier while the other threads clear
    // the working set.  Misaligned caller pointers fall back to plain loads.
    {
        const double* gxl = a.x_lin + (size_t)b * NX * Na;
        const double* gul = a.u_lin + (size_t)b * NU * Na;
        const double* gxr = a.x_ref + (size_t)b * NX * Na;
        const unsigned BX = NX * Na * 8, BU = NU * Na * 8;       // bulk copies need 16-byte sizes and addresses
        const bool aligned = ((((size_t)gxl) | ((size_t)gul) | ((size_t)gxr) | BX | BU) & 15) == 0;
        if (aligned) {
            if (tid == 0) mbar_init(&S.mbar, 1);
            __syncthreads();
            if (tid == 0) {
                mbar_expect_tx(&S.mbar, 2 * BX + BU);
                tma_bulk_g2s(S.xl, gxl, BX, &S.mbar);
                tma_bulk_g2s(S.ul, gul, BU, &S.mbar);
                tma_bulk_g2s(S.xr, gxr, BX, &S.mbar);
            }
        }
        for (int i = tid; i < D::NSLOT; i += NT) S.gi.status[i] = 0;
        if (tid == 0) S.prog = N;
        if constexpr (S_t::RL) {
            for (int i = tid; i < S_t::GR::VL; i += NT) {
                S.gi.x[i] = 0.0; S.gi.g[i] = 0.0; S.gi.rowv[i] = 0.0; S.gi.nvec[i] = 0.0; S.gi.dvec[i] = 0.0;
                S.gi.colk[0][i] = 0.0; S.gi.colk[1][i] = 0.0; S.gi.lam[i] = 0.0; S.gi.cs[i] = 1.0;
            }
        } else {
            for (int i = tid; i < KB_ * RP; i += NT) (&S.gi.nvec[0][0])[i] = 0.0;
            for (int i = tid; i < RP; i += NT) { S.gi.x[i] = 0.0; S.gi.g[i] = 0.0; S.gi.rowv[i] = 0.0; S.gi.zrow[i] = 0.0; S.gi.colk[0][i] = 0.0; S.gi.colk[1][i] = 0.0; }
        }
        if (aligned) {
            mbar_wait(&S.mbar, 0);
        } else {
            for (int i = tid; i < NX * Na; i += NT) { S.xl[i] = gxl[i]; S.xr[i] = gxr[i]; }
            for (int i = tid; i < NU * Na; i += NT) S.ul[i] = gul[i];
        }
    }
    __syncthreads();
    if (PAD && Na < N) {
        // padding steps repeat the last linearisation point (finite Jacobians; they carry no cost and no rows)
        for (int i = NX * Na + tid; i < NX * N; i += NT) { S.xl[i] = S.xl[NX * (Na - 1) + i % NX]; S.xr[i] = 0.0; }
        for (int i = NU * Na + tid; i < NU * N; i += NT) S.ul[i] = S.ul[NU * (Na - 1) + i % NU];
        for (int t = tid; t < D::NROWS; t += NT)
            if (t % N >= Na) S.gi.status[nV + t] = 2;            // disabled: never searched, never reported
        __syncthreads();
    }

    STAGE(0);
    // ---------------------------------------------------------------- linearise + discretise
    if (tid < N) {
        const int k = tid;
        double Ac[NX * NX], Bc_[NX * NU], dc[NX];
        typename Model::Aux aux;
        linearise_step<Model>(P.lin_scheme, S.xl + k * NX, S.ul + k * NU, dt, tr, P, Ac, Bc_, dc, &aux);
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
        C::step_coefs(S.xl + k * NX, S.ul + k * NU, tr, P, S.pc + k * C::NPC, S.g0 + k * C::NG0, &aux);
    } else if (tid == N) {
        C::problem_consts(P, S.cg);
    }
    __syncthreads();

    STAGE(1);
    // ---------------------------------------------------------------- free response + B_bar chains
    // Three things run side by side on different warps:
    //  * warps 0..2: the columns of B_bar (v <- A_k v), packed block-lower-triangular
    //  * warp GW   : the cost Gramian along the horizon, backwards:  W_N = Q_N,
    //                W_s = Q + A_s' W_{s+1} A_s,  G_s = B' W_{s+1}.  With it every entry of
    //                H = 2 (B_bar' Qbar B_bar + Rbar) is ONE short dot product (H stage below)
    //                instead of a sum over the remaining horizon:
    //                H_ij / 2 = e_ci' B' W_{s+1} x^(j)_{s+1},  s = step of the later control i
    //  * warp NW-1 : the free response x_f = A_bar x0 + d_bar
    // With five or more warps the recursion warp runs beside everything else: the other warps
    // ("workers") go on to g / bounds / row norms behind a named barrier of their own and then
    // follow the recursion stage by stage (adjoint rows of J, below); all meet before the tiles
    // are filled.  With four warps the recursion shares the free-response warp and nothing overlaps.
    static_assert(NW >= 4, "stage map: three chain warps + the free-response warp");
    // Small state dimension (kinematic, one matrix entry per lane): ONE recursion warp -- two were measured, no gain (one
    // worker warp fewer, same critical path).  Larger models (dynamic: 49 entries, two passes per phase) with warps to
    // spare run the Gramian and the Riccati recursion on TWO warps side by side (they share nothing but A_s and B).
    constexpr bool OVL = NW >= 5;
    constexpr bool TWO_REC = OVL && NW >= 8 && (NX * NX > 32);
    constexpr int GW = OVL ? 3 : NW - 1;                         // recursion warp (Riccati; both when !TWO_REC)
    constexpr int GW2 = TWO_REC ? GW + 1 : -1;                   // Gramian warp
    constexpr int NREC = OVL ? (TWO_REC ? 2 : 1) : 0;
    constexpr int WNT = NT - 32 * NREC;                          // worker threads
    const bool worker = !OVL || (warp != GW && warp != GW2);
    const int wtid = (OVL && warp > GW) ? tid - 32 * NREC : tid; // worker index
    auto wbar = [&]() {
        if (OVL) asm volatile("bar.sync 1, %0;" ::"r"(WNT) : "memory");
        else __syncthreads();
    };
    if (warp == GW) {
        if constexpr (NX * NX <= 32) horizon_recursions_small<Model, N>(S, P, Na);
        else if constexpr (TWO_REC) horizon_recursions<false, true, Model, N>(S, P, Na);
        else horizon_recursions<true, true, Model, N>(S, P, Na);
    }
    if constexpr (TWO_REC) {
        if (warp == GW2) horizon_recursions<true, false, Model, N>(S, P, Na);
    }
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
    }
    if (tid < 96) {
        for (int t = tid; t < N * NU; t += 96) {
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
                for (int ii = 0; ii < C::NREAL; ++ii) bf_all[ii * D::NPK + D::pk(k, t)] = v[C::real_state(ii)];
                if (LONG && !S_t::BFG) {
#pragma unroll
                    for (int c = 0; c < C::NCR; ++c) S.Bf[c * D::NPK + D::pk(k, t)] = v[C::real_state(C::cons_real(c))];
                }
            }
        }
    }
    if (worker) wbar();        // B_bar and the free response are complete

    STAGE(2);
    // ---------------------------------------------------------------- g, bounds, row norms, cost const
    if (worker) {
        double* e = S.dd;      // tracking error overwrites dd (dead after the free response)
        for (int i = wtid; i < NX * N; i += WNT) e[i] = S.xf[i] - S.xr[i];
        if (S_t::ROWNORMS) {
            // per-step Gram matrix / column sums of the B_bar rows the constraints touch: the squared
            // norm of every row normal then has a closed form (below)
            for (int k = WNT - 1 - wtid; k < N; k += WNT) {      // the last threads: the first ones have the longest g sums
                double G[S_t::NGR], Sm[C::NCR * NU];
#pragma unroll
                for (int i = 0; i < S_t::NGR; ++i) G[i] = 0.0;
#pragma unroll
                for (int i = 0; i < C::NCR * NU; ++i) Sm[i] = 0.0;
                for (int j = 0; j < NU * (k + 1); j += NU) {
#pragma unroll
                    for (int u = 0; u < NU; ++u) {
                        double bc[C::NCR];
#pragma unroll
                        for (int c = 0; c < C::NCR; ++c) bc[c] = S.crow(c)[D::pk(k, j + u)];
                        int gi = 0;
#pragma unroll
                        for (int c = 0; c < C::NCR; ++c) {
                            Sm[c * NU + u] += bc[c];
#pragma unroll
                            for (int c2 = 0; c2 <= c; ++c2, ++gi) G[gi] = fma(bc[c], bc[c2], G[gi]);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < S_t::NGR; ++i) S.gram[k * S_t::NGR + i] = G[i];
#pragma unroll
                for (int i = 0; i < C::NCR * NU; ++i) S.csum[k * C::NCR * NU + i] = Sm[i];
            }
        }
        wbar();
        for (int j = wtid; j < nV; j += WNT) {
            double acc = 0.0;
            if (j < nU) {
                const int sj = j / NU, cj = j - sj * NU;
                for (int k = sj; k < N; ++k) {
#pragma unroll
                    for (int ii = 0; ii < C::NREAL; ++ii) {
                        const int r = C::real_state(ii);
                        const double q = stage_weight(P, k, r, Na);
                        acc += q * bf_all[ii * D::NPK + D::pk(k, j)] * e[k * NX + r];
                    }
#pragma unroll
                    for (int ii = 0; ii < C::NINT; ++ii) {
                        if (C::int_ucol(ii) == cj) {
                            const int r = C::int_state(ii);
                            const double q = stage_weight(P, k, r, Na);
                            acc += q * dt * e[k * NX + r];
                        }
                    }
                }
                acc *= 2.0;
            } else {
                acc = P.R_soft[j - nU];
            }
            S.gi.g[j] = acc;
        }
        if (warp == 0) {
            double acc = 0.0;
            for (int i = lane; i < NX * N; i += 32) {
                const int k = i / NX, r = i - k * NX;
                const double q = stage_weight(P, k, r, Na);
                acc += q * e[i] * e[i];
            }
            acc = warp_sum(acc);
            if (lane == 0) S.scal[0] = acc;
        }
        for (int t = wtid; t < D::NROWS; t += WNT) {
            const int r = t / N, k = t - r * N;
            double lo, up;
            C::row_bounds(r, S.xf + k * NX, S.xl + k * NX, S.ul + k * NU, S.pc + k * C::NPC, S.g0 + k * C::NG0, S.cg, P, lo, up);
            S.rlo[t] = lo;
            S.rup[t] = up;
            // squared norm of the row normal (u part + slack entry); only a scale for the
            // linear-dependence threshold.  With v_j = sum_c a_c B_c[k][j] + b_u(j) + [step(j) = k] cu_u(j):
            //   |v|^2 = a'G a + 2 sum_u b_u (a'S_u) + (k+1) sum_u b_u^2 + sum_u cu_u (2 (a'B[.,2k+u] + b_u) + cu_u)
            double n2 = 1.0;
            if (S_t::ROWNORMS) {
                const double* pc = S.pc + k * C::NPC;
                double ac[C::NCR], bu[NU];
#pragma unroll
                for (int c = 0; c < C::NCR; ++c) ac[c] = C::row_coef(r, c, pc, S.cg);
#pragma unroll
                for (int u = 0; u < NU; ++u) bu[u] = 0.0;
#pragma unroll
                for (int c = 0; c < C::NINT; ++c) bu[C::int_ucol(c)] += C::row_coef(r, C::NCR + c, pc, S.cg) * dt;
                n2 = (C::row_slack(r) >= 0) ? 1.0 : 0.0;
                int gi = 0;
#pragma unroll
                for (int c = 0; c < C::NCR; ++c) {
#pragma unroll
                    for (int c2 = 0; c2 <= c; ++c2, ++gi) n2 += (c2 == c ? 1.0 : 2.0) * ac[c] * ac[c2] * S.gram[k * S_t::NGR + gi];
                }
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    double aS = 0.0, aB = 0.0;
#pragma unroll
                    for (int c = 0; c < C::NCR; ++c) {
                        aS += ac[c] * S.csum[(k * C::NCR + c) * NU + u];
                        aB += ac[c] * S.crow(c)[D::pk(k, NU * k + u)];
                    }
                    const double cu = C::row_ucoef(r, u, pc, S.cg);
                    n2 += bu[u] * (2.0 * aS + (double)(k + 1) * bu[u]) + cu * (2.0 * (aB + bu[u]) + cu);
                }
            }
            S.rn2[t] = n2;
        }

        // Rows of J, following the recursion: row t = (k, cu) is  e_cu' K_k Acl_{k-1} .. Acl_{m+1} B Lambda_m^(-T/2)
        // at the columns of stage m < k (Acl = A + B K), and Lambda_k^(-T/2) on the diagonal block (the
        // 2 x 2 factor Lambda_m^(-T/2) / sqrt 2 of every column pair is applied when the tiles are filled).  The
        // adjoint  lam <- lam Acl_m  runs in the SAME direction as the Riccati recursion, so each row
        // thread advances one stage as soon as that stage is published.  Staged packed (like B_bar)
        // in the region that later holds the packed H.
        {
            static_assert(NU * D::NPK <= D::HP, "J staging must fit the packed-H region");
            static_assert(WNT >= nU, "one worker thread per row of J");
            double* Jst = S.gi.hp();
            const int t = wtid - (WNT - nU);
            const int wfirst = (WNT - nU) >> 5;                 // first worker warp that owns rows
            if ((wtid >> 5) >= wfirst) {
                const int k = t >= 0 ? t / NU : N, cu = t >= 0 ? t - (t / NU) * NU : 0;   // k = N: no row (padding lanes)
                double lamv[NX];
#pragma unroll
                for (int l = 0; l < NX; ++l) lamv[l] = 0.0;
                for (int mm = N - 1; mm >= 0; --mm) {
                    while (*(volatile int*)&S.prog > mm) __nanosleep(100);
                    __threadfence_block();
                    if (mm == k) {
#pragma unroll
                        for (int c = 0; c < NU; ++c) Jst[cu * D::NPK + D::pk(k, NU * k + c)] = (c == cu) ? 1.0 : 0.0;
                        if (k > 0) {
#pragma unroll
                            for (int l = 0; l < NX; ++l) lamv[l] = S.Kc[(k * NU + cu) * NX + l];
                        }
                    } else if (mm < k && k < N) {
                        double bv[NU];
#pragma unroll
                        for (int r = 0; r < NU; ++r) {
                            double acc = 0.0;
#pragma unroll
                            for (int l = 0; l < NX; ++l) acc = fma(lamv[l], S.B1[l * NU + r], acc);
                            bv[r] = acc;
                        }
#pragma unroll
                        for (int c = 0; c < NU; ++c) Jst[cu * D::NPK + D::pk(k, NU * mm + c)] = bv[c];    // x Lambda_m^(-T/2) at the tile fill
                        if (mm > 0) {
                            double ln[NX];
#pragma unroll
                            for (int l = 0; l < NX; ++l) {
                                double acc = (l >= C::NREAL) ? lamv[l] : 0.0;
#pragma unroll
                                for (int jr = 0; jr < C::NREAL; ++jr) acc = fma(lamv[jr], S.Ad[(mm * C::NREAL + jr) * NX + l], acc);
#pragma unroll
                                for (int r = 0; r < NU; ++r) acc = fma(bv[r], S.Kc[(mm * NU + r) * NX + l], acc);
                                ln[l] = acc;
                            }
#pragma unroll
                            for (int l = 0; l < NX; ++l) lamv[l] = ln[l];
                        }
                    }
                }
            }
        }
    }

    __syncthreads();           // g, bounds, J staging complete
    STAGE(3);
    // ---------------------------------------------------------------- operator tiles, packed H
    // M = [ e_{nU}, .., e_{nU+NS-1} | J ]: the NS flat (zero-curvature) slack variables start with their
    // lower bound in the working set (q = NS, lam = R_soft: dual feasible), J (J'HJ = I) from the staging.
    constexpr bool RL = S_t::RL;
    using GR = typename S_t::GR;
    using Ops = std::conditional_t<RL, RlOps<GR, typename S_t::Gi_t>, GiOps<G, typename S_t::Gi_t>>;
    typename S_t::Gi_t& Q = S.gi;
    std::conditional_t<RL, RlTile<GR>, GiTile<G>> m;
    double lam[CS];
    int q = NS, ybuf = 0;
    double g_w = 0.0;
    if constexpr (RL) {
        // Row-lane tiles: thread (warp, lane) holds rows lane + 32 s, columns warp * CPW + c.  Integrator coordinates:
        // the operator the loop works on is T M (rows (k, c) = dt * the prefix sum of rows (0..k, c) of M); every
        // later update is a column operation, so it commutes with T and the loop produces T z, T x.  The prefix
        // runs over the staged rows in shared memory (it commutes with the 2 x 2 column factors applied below).
        static_assert(NT >= nU && NU == 2, "one thread per control for the g stencil; two interleaved channels");
        double* Jst = S.gi.hp();
        auto mval = [&](int i, int jj) -> double {        // entry (i, jj) of [e_slack | J] from the staging
            if (jj < NS) return (i == nU + jj) ? 1.0 : 0.0;
            if (jj >= nV || i >= nU) return 0.0;
            const int j = jj - NS, sj = j / NU, cj = j - sj * NU, si = i / NU, ci = i - si * NU;
            if (si < sj) return 0.0;
            const double* jr = Jst + ci * D::NPK + D::pk(si, sj * NU);
            double v = 0.0;
#pragma unroll
            for (int r2 = 0; r2 < NU; ++r2) v = fma(jr[r2], S.Wi[sj * NU * NU + r2 * NU + cj], v);
            return v;
        };
        if (a.dbg_M) {                                   // debug tap: the operator in control coordinates
            double* gM = a.dbg_M + (size_t)b * nV * nV;
            for (int t = tid; t < nV * nV; t += NT) gM[t] = mval(t % nV, t / nV);
        }
        if (a.dbg_g) {
            double* gg = a.dbg_g + (size_t)b * nV;
            for (int t = tid; t < nV; t += NT) gg[t] = S.gi.g[t];
        }
        if (tid < nU) g_w = (S.gi.g[tid] - (tid + NU < nU ? S.gi.g[tid + NU] : 0.0)) * S.gi.idt;   // T^-T g
        if (tid < GR::VL) S.gi.lam[tid] = (tid < NS) ? fabs(S.gi.g[nU + tid]) : 0.0;
        __syncthreads();
        for (int task = tid; task < NU * nU; task += NT) {
            const int ci = task / nU, col = task - ci * nU;
            double acc = 0.0;
            for (int si = col / NU; si < N; ++si) {
                double* pj = Jst + ci * D::NPK + D::pk(si, col);
                acc += *pj;
                *pj = acc * dt;
            }
        }
        if (tid < nU) S.gi.g[tid] = g_w;
        __syncthreads();
#pragma unroll
        for (int sl = 0; sl < GR::RS; ++sl) {
#pragma unroll
            for (int c = 0; c < GR::CPW; ++c) m(sl, c) = mval(lane + 32 * sl, warp * GR::CPW + c);
        }
        __syncthreads();           // staging consumed: the region becomes the packed H
    } else {
    {
        const double* Jst = S.gi.hp();
        if (LONG) m.attach(S.Msm);
#pragma unroll
        for (int s = 0; s < CS; ++s) {
            const int jj = lane + 32 * s, j = jj - NS;
            const int sj = (j >= 0) ? j / NU : 0;
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const int i = row0 + r;
                double v = 0.0;
                if (jj < NS) v = (i == nU + jj) ? 1.0 : 0.0;
                else if (jj < nV && i < nU) {
                    const int si = i / NU, ci = i - si * NU;
                    if (si >= sj) {                 // raw row entries of the column pair (sj, .) times Lambda_sj^(-T/2)/sqrt 2
                        const int cj = j - sj * NU;
                        const double* jr = Jst + ci * D::NPK + D::pk(si, sj * NU);
                        v = 0.0;
#pragma unroll
                        for (int r2 = 0; r2 < NU; ++r2) v = fma(jr[r2], S.Wi[sj * NU * NU + r2 * NU + cj], v);
                    }
                }
                m(r, s) = v;
            }
            lam[s] = (jj < NS) ? fabs(S.gi.g[nU + jj]) : 0.0;
        }
    }
    if (a.dbg_M) {
        double* gM = a.dbg_M + (size_t)b * nV * nV;
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int i = row0 + r;
#pragma unroll
            for (int s = 0; s < CS; ++s) {
                const int j = lane + 32 * s;
                if (i < nV && j < nV) gM[(size_t)j * nV + i] = m(r, s);
            }
        }
    }
    // Integrator coordinates (C::WSPACE) with the column-lane tiles (long horizons): T M by a prefix inside the
    // thread's rows, then the carry of the warps below (published through the ypart scratch).  g becomes T^-T g.
    if constexpr (C::WSPACE && !RL) {
        static_assert(NT >= nU && NU == 2, "one thread per control for the g stencil; two interleaved channels");
        if (a.dbg_g) {
            double* gg = a.dbg_g + (size_t)b * nV;
            for (int t = tid; t < nV; t += NT) gg[t] = S.gi.g[t];
        }
        if (tid < nU) g_w = (S.gi.g[tid] - (tid + NU < nU ? S.gi.g[tid + NU] : 0.0)) * S.gi.idt;
        double cv[2][CS];
#pragma unroll
        for (int s = 0; s < CS; ++s) { cv[0][s] = 0.0; cv[1][s] = 0.0; }
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int i = row0 + r;
            if (i < nU) {                               // warp-uniform
#pragma unroll
                for (int s = 0; s < CS; ++s) {
                    if (r >= 2) m(r, s) += m(r >= 2 ? r - 2 : 0, s);
                    if (i & 1) cv[1][s] = m(r, s); else cv[0][s] = m(r, s);
                }
            }
        }
#pragma unroll
        for (int s = 0; s < CS; ++s) {
            S.gi.ypart[0][warp][lane + 32 * s] = cv[0][s];
            S.gi.ypart[1][warp][lane + 32 * s] = cv[1][s];
        }
    }
    __syncthreads();           // staging consumed: the region becomes the packed H
    if constexpr (C::WSPACE && !RL) {
        if (tid < nU) S.gi.g[tid] = g_w;
        double c0[CS], c1[CS];
#pragma unroll
        for (int s = 0; s < CS; ++s) { c0[s] = 0.0; c1[s] = 0.0; }
        for (int w = 0; w < warp; ++w) {
#pragma unroll
            for (int s = 0; s < CS; ++s) {
                c0[s] += S.gi.ypart[0][w][lane + 32 * s];
                c1[s] += S.gi.ypart[1][w][lane + 32 * s];
            }
        }
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int i = row0 + r;
            if (i < nU) {
#pragma unroll
                for (int s = 0; s < CS; ++s) m(r, s) = (m(r, s) + ((i & 1) ? c1[s] : c0[s])) * dt;
            }
        }
    }
    }
    // generate_qp.m:29  H = 2 (B' Qbar B + Rbar), packed lower triangle for the symv's (drops, refresh,
    // objective).  Entry (i, j), i >= j, i the later control at step si:
    //   H_ij = 2 G_si[c_i] . x^(j)_{si+1}  (+ 2 R on the diagonal)
    // x^(j)_{si+1} = column j of B_bar at step si: packed rows for the real states, exactly dt for the
    // integrator state its control drives.  The slack diagonal gets flat_eps.
    {
        static_assert(C::real_state(0) == 0 && C::real_state(C::NREAL - 1) == C::NREAL - 1, "real states come first");
        double* gH = a.dbg_H ? a.dbg_H + (size_t)b * nV * nV : nullptr;
        // one warp per row i (rows dealt from the longest down, so the warps finish together), lanes over
        // the columns j <= i: row quantities are warp-uniform, packed loads and stores are contiguous
        for (int ir = warp; ir < nV; ir += NW)
        for (int i = nV - 1 - ir, j = lane; j <= i; j += 32) {
            double h = 0.0;
            if (i < nU) {
                const int si = i / NU, ci = i - si * NU, cj = j % NU;
                const double* Gr = S.Gs + (si * NU + ci) * NX;
                double acc = 0.0;
#pragma unroll
                for (int ii = 0; ii < C::NREAL; ++ii) acc = fma(Gr[ii], bf_all[ii * D::NPK + D::pk(si, j)], acc);
#pragma unroll
                for (int ii = 0; ii < C::NINT; ++ii)
                    if (C::int_ucol(ii) == cj) acc = fma(Gr[C::int_state(ii)], dt, acc);
                h = 2.0 * acc;
                if (i == j) h += 2.0 * P.R[ci];
            }
            if (gH) { gH[(size_t)j * nV + i] = h; gH[(size_t)i * nV + j] = h; }
            const double hv = (i >= nU) ? (i == j ? P.flat_eps : 0.0) : h;
            if (LONG) {                             // full symmetric matrix in the L2 slab: coalesced symv's
                S.gi.hp()[(size_t)i * nV + j] = hv;
                S.gi.hp()[(size_t)j * nV + i] = hv;
            } else {
                S.gi.hp()[D::hp(i, j)] = hv;
            }
        }
    }
    if (!C::WSPACE && a.dbg_g) {
        double* gg = a.dbg_g + (size_t)b * nV;
        for (int t = tid; t < nV; t += NT) gg[t] = S.gi.g[t];
    }
    STAGE(4);
    if (tid < NS) {
        Q.act[tid] = (nU + tid) * 2;                       // slack lower bounds in the working set,
        Q.status[nU + tid] = -1;                           // multiplier R_soft (dual feasible start)
    }
    __syncthreads();
    if constexpr (C::WSPACE) {
        // packed B_bar rows -> integrator coordinates: (B T^-1)[k][j] = (B[k][j] - B[k][j+2]) / dt, in place, ascending
        // j (the entry two to the right is still the original one).  One task per (row, channel); tasks of steps k
        // and N-1-k are paired so that every thread has N+1 entries.
        static_assert(N % 2 == 0, "step pairs");
        const double idt = S.gi.idt;
        constexpr int NROWSETS = C::NREAL + ((LONG && !S_t::BFG) ? C::NCR : 0);
        for (int task = tid; task < NROWSETS * (N / 2) * NU; task += NT) {
            const int c = task % NU, t2 = task / NU, g2 = t2 % (N / 2), rs = t2 / (N / 2);
            double* rowbase = (rs < C::NREAL) ? bf_all + rs * D::NPK : S.Bf + (rs - C::NREAL) * D::NPK;
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const int k = half ? N - 1 - g2 : g2;
                double* row = rowbase + D::pk(k, 0);
                double cur = row[c];
                for (int j = c; j < NU * (k + 1); j += NU) {
                    const double nxt = (j + NU < NU * (k + 1)) ? row[j + NU] : 0.0;
                    row[j] = (cur - nxt) * idt;
                    cur = nxt;
                }
            }
        }
        __syncthreads();
    }
    STAGE(5);
    if constexpr (RL) Ops::initial_point(Q, m, q, nU, nV);
    else Ops::initial_point(Q, m, ybuf, q, nU, nV);        // x_u = -J J' g, slacks at 0
    STAGE(6);
    const MpcProb<Model, N, NW_, KB_, CSR_> prob{S, P, dt};
    const int iter_cap = P.max_iter > 0 ? P.max_iter : 5 * (NU * Na + NS + C::n_ref_rows(Na));
    GiStats st;
    if constexpr (RL) st = Ops::solve(prob, Q, m, q, nV, P.feas_tol, iter_cap);
    else st = Ops::solve(prob, Q, m, lam, q, ybuf, nV, P.feas_tol, iter_cap);
    const int iters = st.iters, exitflag = st.exitflag, n_add = st.n_add, n_drop = st.n_drop, n_refresh = st.n_refresh;

    STAGE(7);
    // ---------------------------------------------------------------- outputs
    // fval = 1/2 x'Hx + g'x + const (ltvmpc_*_curvilinear.m:60); H without the flat_eps entries
    {
        double f;
        if constexpr (RL) f = Ops::objective(Q, nV);
        else f = Ops::objective(Q, nV, nullptr);
        if (tid == 0) {
            for (int j = nU; j < nV; ++j) f -= 0.5 * P.flat_eps * Q.x[j] * Q.x[j];
            a.fval[b] = f + S.scal[0];
            // a non-finite objective means a non-finite iterate: qpOASES's "internal error", never "solved"
            a.exitflag[b] = (exitflag == GI_EXIT_SOLVED && !isfinite(f)) ? -1 : exitflag;
            if (a.iters) a.iters[b] = iters;
            if (a.counters) {
                atomicAdd(a.counters + 0, (unsigned long long)n_add);
                atomicAdd(a.counters + 1, (unsigned long long)n_drop);
                atomicAdd(a.counters + 2, (unsigned long long)n_refresh);
            }
        }
    }
    for (int j = tid; j < NU * Na; j += NT)
        a.u_opt[(size_t)b * NU * Na + j] = C::WSPACE ? (S.gi.x[j] - (j >= NU ? S.gi.x[j - NU] : 0.0)) * S.gi.idt : S.gi.x[j];
    for (int j = tid; j < NS; j += NT) a.slack_opt[(size_t)b * NS + j] = S.gi.x[nU + j];
    // x_opt = A_bar x0 + B_bar u + d_bar = xf + B_bar u  (ltvmpc_*_curvilinear.m:58)
    {
        double* gxo = a.x_opt + (size_t)b * NX * Na;
        constexpr int NRR = C::NREAL * N;
        for (int base = 0; base < NRR; base += NT / 4) {
            const int rid = base + (tid >> 2), part = tid & 3;
            double acc = 0.0;
            int c = 0, k = 0;
            if (rid < NRR) {
                c = rid / N; k = rid - c * N;
                const double* row = bf_all + c * D::NPK + D::pk(k, 0);
                const int len = NU * (k + 1);
                const int ch = (len + 3) >> 2;
                const int j0 = part * ch, j1 = (j0 + ch < len) ? j0 + ch : len;
                for (int j = j0; j < j1; ++j) acc += row[j] * S.gi.x[j];
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            if (rid < NRR && part == 0 && k < Na) gxo[k * NX + C::real_state(c)] = S.xf[k * NX + C::real_state(c)] + acc;
        }
        for (int t = tid; t < C::NINT * N; t += NT) {
            const int ci = t / N, k = t - ci * N;
            const int uc = C::int_ucol(ci), rs = C::int_state(ci);
            double acc = 0.0;
            if (C::WSPACE) acc = S.gi.x[NU * k + uc];          // the integrator coordinate IS the perturbation of this state
            else {
                for (int i = 0; i <= k; ++i) acc += S.gi.x[NU * i + uc];
                acc *= dt;
            }
            if (k < Na) gxo[k * NX + rs] = S.xf[k * NX + rs] + acc;
        }
    }
    if (a.wsB) {
        const int nUa = NU * Na;               // the caller's variable order: controls of the Na steps, then the slacks
        for (int j = tid; j < nUa + NS; j += NT) a.wsB[(size_t)b * (nUa + NS) + j] = S.gi.status[j < nUa ? j : nU + (j - nUa)];
    }
    if (a.wsC) {
        int8_t* w = a.wsC + (size_t)b * C::n_ref_rows(Na);
        for (int j = tid; j < C::n_ref_rows(Na); j += NT) w[j] = 0;
        __syncthreads();
        for (int j = tid; j < q; j += NT) {
            const int code = S.gi.act[j], slot = code >> 1, side = (code & 1) ? +1 : -1;
            if (slot >= nV) {
                const int rr = slot - nV, r = rr / N, k = rr - r * N;
                w[C::ref_row(r, k, side, Na)] = (int8_t)side;
            }
        }
    }
    STAGE(8);
}

}  // namespace fsae
