// Dual active-set core, ROW-LANE operator layout ("RL"): the second generation of gi_core.cuh for problems whose
// constraint normals are SPARSE (the kinematic LTV-MPC QPs in integrator coordinates, fused_v2.cuh).
//
// Same method as gi_core.cuh -- Goldfarb-Idnani in operator form, M = [K1 | J2], same pivot rule, step lengths,
// refresh and exit flags -- but the nV x nV operator is tiled the other way round:
//     lane l  owns ROWS    l, l+32, l+64, ...        (RS = ceil(nV/32) row slots per lane)
//     warp w  owns COLUMNS w*CPW .. w*CPW+CPW-1      (CPW = ceil(nV/NW))
// i.e. every thread holds an RS x CPW tile (3 x 14 doubles for nV = 81 with 6 warps).  With that layout
//   * y = M'n for a normal with <= 3 entries is a combination of <= 3 rows of M: the lanes that own those rows
//     scale their CPW entries and hand them to the warp through shared memory -- no product, no reduction, no
//     block barrier, and every warp needs y only for ITS columns;
//   * z = J2 y2 is 42 FMAs per thread + one cross-warp sum (the only block-wide exchange of an iteration; the
//     partial step-length quantities d2 = |y2|^2 and min lam_j / y_j ride on the same barrier);
//   * the rank-1 / Householder update needs the row scalars k_i, w_i only for the thread's RS rows (not for all
//     rows of a warp), and warps whose columns lie entirely left or right of column q do one FMA per element.
// The in-warp reduce-scatter of the column-lane layout (15 shuffle rounds per iteration) is gone; it survives only
// for DENSE vectors (the rare dense normals, drops, the refresh), as the reduction of y = M'v over the lanes.
// An add iteration has three block barriers: arg-min of the search, the exchange above, end of the step.
#pragma once
#include "gi_core.cuh"

namespace fsae {

template <int NVMAX, int NW_>
struct RlCfg {
    static constexpr int NW = NW_, NT = 32 * NW_;
    static constexpr int KB = 1;
    static constexpr int CPW = (NVMAX + NW - 1) / NW;        // columns per warp
    static constexpr int CPW2 = (CPW + 1) & ~1;              // ... padded to whole 16-byte pairs
    static constexpr int CP = CPW * NW;                      // padded columns
    static constexpr int RS = (NVMAX + 31) / 32;             // row slots per lane
    static constexpr int RP = 32 * RS;                       // padded rows
    static constexpr int VL = RP > CP ? RP : CP;             // length of every row- or column-indexed vector
    static constexpr int RH = CPW <= 8 ? 8 : (CPW <= 16 ? 16 : 32);   // reduce-scatter width (dense y = M'v)
    static constexpr int SP = 3;                             // partial sums per row of a packed symv
    static constexpr int HP = NVMAX * (NVMAX + 1) / 2;       // packed lower triangle
    static_assert(CPW <= 32, "one lane per column of a warp");
    __host__ __device__ static constexpr int hp(int i, int j) { return i * (i + 1) / 2 + j; }   // i >= j
};

// Shared-memory working set.  WSP as in GiSm: the first WSP variables are integrator coordinates, the packed H
// stays in control coordinates and every H v is wrapped in the two difference stencils.
// SR >= 0: row SR of the operator (the slack every soft row shares) is mirrored per warp in srow, so that a normal
// "two row entries + a multiple of row SR" needs one publishing pass only.
template <class G, int NSLOT, int WSP_ = 0, int SR_ = -1>
struct RlSm {
    static constexpr int WSP = WSP_, SR = SR_;
    static constexpr bool HFULL = false;
    double idt;
    alignas(16) double x[G::VL];
    double g[G::VL];
    double Hp[G::HP];
    __device__ __forceinline__ double* hp() { return Hp; }
    double lam[G::VL];                 // multiplier of working-set column j
    double cs[G::VL];                  // column scales: the true column j of the operator is cs[j] * (tile column j)
    double colk[2][G::VL];             // column broadcasts (column q / drop column / moved column)
    double rowv[G::VL];                // H v
    double nvec[G::VL];                // dense normal
    double dvec[G::VL];                // T^-1 v scratch of the symv
    double wpart[G::SP][G::VL];        // symv partials
    alignas(16) double ysm[G::NW][4][G::CPW2];   // per warp: row parts / masked copies of y for its columns
    alignas(16) double srow[G::NW][G::CPW2];     // row SR of the operator, this warp's columns (kept current by its owner lane)
    double zpart[G::NW][G::RP];        // cross-warp partial sums of z = M y
    double red_d2[G::NW];
    double red_yq;
    unsigned long long red_t1[G::NW];
    int red_l[G::NW];
    unsigned long long red_key[2][G::NW];
    int red_idx[2][G::NW];
    unsigned long long red_ymin[G::NW];  // refresh: most negative recomputed multiplier per warp
    unsigned long long red_ymax[G::NW];
    int red_lmin[G::NW];
    double red_val[G::NW];
    int act[G::VL];                    // slot*2 + (side > 0) of working-set column j
    int8_t status[NSLOT + 8];          // -1 / 0 / +1 per slot
};

template <class G>
struct RlTile {
    double v[G::RS][G::CPW];
    __device__ __forceinline__ double& operator()(int s, int c) { return v[s][c]; }
    __device__ __forceinline__ const double& operator()(int s, int c) const { return v[s][c]; }
};

template <class G, class SM>
struct RlOps {
    static constexpr int NW = G::NW, NT = G::NT, CPW = G::CPW, CPW2 = G::CPW2, RS = G::RS, RH = G::RH;
    using Tile = RlTile<G>;

    // all lanes: the warp's CPW2 published values of buffer b
    __device__ __forceinline__ static void load_y(const SM& S, int b, double (&y)[CPW]) {
        const int warp = threadIdx.x >> 5;
        const double2* p = reinterpret_cast<const double2*>(S.ysm[warp][b]);
#pragma unroll
        for (int h = 0; h < CPW2 / 2; ++h) {
            const double2 t = p[h];
            y[2 * h] = t.x;
            if (2 * h + 1 < CPW) y[2 * h + 1] = t.y;
        }
    }

    // y = M'v for this warp's columns, v a full-length vector in shared memory (visible block-wide).
    // On return: y[c] in every lane, ymine = y[lane] (lanes >= CPW: 0), and ysm[warp][3] holds y.
    __device__ __forceinline__ static void y_dense(SM& S, const Tile& m, const double* v, double (&y)[CPW], double& ymine) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        double vr[RS];
#pragma unroll
        for (int s = 0; s < RS; ++s) vr[s] = v[lane + 32 * s];
        double p[RH];
#pragma unroll
        for (int c = 0; c < RH; ++c) p[c] = 0.0;
#pragma unroll
        for (int s = 0; s < RS; ++s) {
#pragma unroll
            for (int c = 0; c < CPW; ++c) p[c] = fma(m(s, c), vr[s], p[c]);
        }
        const double r = warp_reduce_scatter<RH>(p);
        constexpr int SH = (RH == 32) ? 0 : (RH == 16 ? 1 : 2);
        const int cc = lane >> SH;
        if ((lane & ((1 << SH) - 1)) == 0 && cc < CPW2) S.ysm[warp][3][cc] = (cc < CPW) ? r : 0.0;
        __syncwarp();
        load_y(S, 3, y);
        ymine = (lane < CPW) ? S.ysm[warp][3][lane] : 0.0;
    }

    // Sparse normal, split for the lanes (once per pivot):  n = sum_{e < cnt} cf[e] e_{idx[e]} + cfs e_{SR}.  Every lane owns
    // at most one of the (<= 2) row entries; smask = the row slots that hold one.
    struct Sparse {
        int cnt, smask;
        int own, mys;            // this lane's entry (or -1) and its row slot
        double mycf, cfs;
    };
    __device__ __forceinline__ static Sparse sparse_split(const SpN& n) {
        const int lane = threadIdx.x & 31;
        Sparse sp;
        int cnt = n.cnt;
        sp.cfs = 0.0;
        if constexpr (SM::SR >= 0) {
            int drop = 0;
#pragma unroll
            for (int e = 0; e < 3; ++e)
                if (e == cnt - 1 && n.idx(e) == SM::SR) { sp.cfs = n.cf(e); drop = 1; }
            cnt -= drop;
        }
        sp.cnt = cnt;
        sp.smask = 0;
        sp.own = -1;
        sp.mys = 0;
        sp.mycf = 0.0;
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            if (e < cnt) {
                sp.smask |= 1 << (n.idx(e) >> 5);
                if ((n.idx(e) & 31) == lane) { sp.own = e; sp.mys = n.idx(e) >> 5; sp.mycf = n.cf(e); }
            }
        }
        return sp;
    }
    // the owner lane of row SR mirrors its entries (after every change of the tiles)
    __device__ __forceinline__ static void publish_srow(SM& S, const Tile& m) {
        if constexpr (SM::SR >= 0) {
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            if (lane == (SM::SR & 31)) {
                double2* dst = reinterpret_cast<double2*>(S.srow[warp]);
#pragma unroll
                for (int h = 0; h < CPW2 / 2; ++h) {
                    double2 t;
                    t.x = m(SM::SR >> 5, 2 * h);
                    t.y = (2 * h + 1 < CPW) ? m(SM::SR >> 5, (2 * h + 1 < CPW) ? 2 * h + 1 : 0) : 0.0;
                    dst[h] = t;
                }
            }
        }
    }
    // y = M'n for a sparse normal.  The lanes that own the row entries scale their CPW entries of that row and publish
    // them (one pass per row slot involved, usually one); lane c sums the parts of column c and the mirrored row SR.
    __device__ __forceinline__ static void y_sparse(SM& S, const Tile& m, const Sparse& sp, double (&y)[CPW], double& ymine) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int s = 0; s < RS; ++s) {
            if (sp.smask & (1 << s)) {                          // warp-uniform
                if (sp.own >= 0 && sp.mys == s) {
                    double2* dst = reinterpret_cast<double2*>(S.ysm[warp][sp.own]);
#pragma unroll
                    for (int h = 0; h < CPW2 / 2; ++h) {
                        double2 t;
                        t.x = sp.mycf * m(s, 2 * h);
                        t.y = (2 * h + 1 < CPW) ? sp.mycf * m(s, (2 * h + 1 < CPW) ? 2 * h + 1 : 0) : 0.0;
                        dst[h] = t;
                    }
                }
            }
        }
        __syncwarp();
        double acc = 0.0;
        if (lane < CPW2) {
            if (SM::SR >= 0) acc = sp.cfs * S.srow[warp][lane];
            if (sp.cnt > 0) acc += S.ysm[warp][0][lane];
            if (sp.cnt > 1) acc += S.ysm[warp][1][lane];
            if (sp.cnt > 2) acc += S.ysm[warp][2][lane];
            S.ysm[warp][3][lane] = acc;
        }
        __syncwarp();
        load_y(S, 3, y);
        ymine = (lane < CPW) ? acc : 0.0;
    }

    // Partial sums of z = sum_{j >= q0} M[:, j] y_j over this warp's columns -> S.zpart[warp][.]; warps whose columns
    // all lie left of q0 write nothing (z_sum skips them).  The MIXED warp (q0 inside its columns) uses a masked copy
    // of y that its column lanes publish (ymine = y of column col0 + lane): no selects on register-resident vectors.
    __device__ __forceinline__ static void z_part(SM& S, const Tile& m, const double (&y)[CPW], double ymine, int q0) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int c0 = q0 - warp * CPW;             // first local column that takes part (warp-uniform)
        if (c0 >= CPW) return;
        double zp[RS];
#pragma unroll
        for (int s = 0; s < RS; ++s) zp[s] = 0.0;
        if (c0 <= 0) {
#pragma unroll
            for (int c = 0; c < CPW; ++c) {
#pragma unroll
                for (int s = 0; s < RS; ++s) zp[s] = fma(m(s, c), y[c], zp[s]);
            }
        } else {
            if (lane < CPW2) S.ysm[warp][1][lane] = (lane >= c0 && lane < CPW) ? ymine : 0.0;
            __syncwarp();
            double yj[CPW];
            load_y(S, 1, yj);
#pragma unroll
            for (int c = 0; c < CPW; ++c) {
#pragma unroll
                for (int s = 0; s < RS; ++s) zp[s] = fma(m(s, c), yj[c], zp[s]);
            }
        }
#pragma unroll
        for (int s = 0; s < RS; ++s) S.zpart[warp][lane + 32 * s] = zp[s];
    }
    // after the barrier: z for this thread's rows (identical in every warp)
    __device__ __forceinline__ static void z_sum(const SM& S, int q0, double (&z)[RS]) {
        const int lane = threadIdx.x & 31;
        const int w0 = q0 / CPW;
#pragma unroll
        for (int s = 0; s < RS; ++s) z[s] = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            if (w >= w0) {
#pragma unroll
                for (int s = 0; s < RS; ++s) z[s] += S.zpart[w][lane + 32 * s];
            }
        }
    }

    // store column c (runtime, warp-uniform) of this warp's tile to a row-indexed vector / overwrite it: a jump
    // table over the CPW compile-time column indices (registers cannot be indexed at run time)
    template <int C>
    __device__ __forceinline__ static void col_store_c(const Tile& m, double* dst) {
        if constexpr (C < CPW) {
            const int lane = threadIdx.x & 31;
#pragma unroll
            for (int s = 0; s < RS; ++s) dst[lane + 32 * s] = m(s, C);
        }
    }
    template <int C>
    __device__ __forceinline__ static void col_set_c(Tile& m, const double (&v)[RS]) {
        if constexpr (C < CPW) {
#pragma unroll
            for (int s = 0; s < RS; ++s) m(s, C) = v[s];
        }
    }
#define FSAE_RL_COLSWITCH(fn, ...) \
    switch (c) { \
        case 0: fn<0>(__VA_ARGS__); break; \
        case 1: fn<1>(__VA_ARGS__); break; \
        case 2: fn<2>(__VA_ARGS__); break; \
        case 3: fn<3>(__VA_ARGS__); break; \
        case 4: fn<4>(__VA_ARGS__); break; \
        case 5: fn<5>(__VA_ARGS__); break; \
        case 6: fn<6>(__VA_ARGS__); break; \
        case 7: fn<7>(__VA_ARGS__); break; \
        case 8: fn<8>(__VA_ARGS__); break; \
        case 9: fn<9>(__VA_ARGS__); break; \
        case 10: fn<10>(__VA_ARGS__); break; \
        case 11: fn<11>(__VA_ARGS__); break; \
        case 12: fn<12>(__VA_ARGS__); break; \
        case 13: fn<13>(__VA_ARGS__); break; \
        case 14: fn<14>(__VA_ARGS__); break; \
        case 15: fn<15>(__VA_ARGS__); break; \
        case 16: fn<16>(__VA_ARGS__); break; \
        case 17: fn<17>(__VA_ARGS__); break; \
        case 18: fn<18>(__VA_ARGS__); break; \
        case 19: fn<19>(__VA_ARGS__); break; \
        case 20: fn<20>(__VA_ARGS__); break; \
        case 21: fn<21>(__VA_ARGS__); break; \
        case 22: fn<22>(__VA_ARGS__); break; \
        case 23: fn<23>(__VA_ARGS__); break; \
        case 24: fn<24>(__VA_ARGS__); break; \
        case 25: fn<25>(__VA_ARGS__); break; \
        case 26: fn<26>(__VA_ARGS__); break; \
        case 27: fn<27>(__VA_ARGS__); break; \
        case 28: fn<28>(__VA_ARGS__); break; \
        case 29: fn<29>(__VA_ARGS__); break; \
        case 30: fn<30>(__VA_ARGS__); break; \
        case 31: fn<31>(__VA_ARGS__); break; \
        default: break; \
    }
    __device__ __forceinline__ static void col_store(const Tile& m, int c, double* dst) { FSAE_RL_COLSWITCH(col_store_c, m, dst) }
    __device__ __forceinline__ static void col_set(Tile& m, int c, const double (&v)[RS]) { FSAE_RL_COLSWITCH(col_set_c, m, v) }
#undef FSAE_RL_COLSWITCH

    // rowv = H v (+ addv), all rows, visible block-wide on return (barriers inside).  v full length in shared memory.
    __device__ __forceinline__ static void symv(SM& S, const double* v, const double* addv, int nV) {
        const int tid = threadIdx.x;
        constexpr int SP = G::SP;
        const int CH = (nV + SP - 1) / SP;
        if constexpr (SM::WSP > 0) {
            for (int i = tid; i < nV; i += NT) S.dvec[i] = (i < SM::WSP) ? (v[i] - (i >= 2 ? v[i - 2] : 0.0)) * S.idt : v[i];
            __syncthreads();
            v = S.dvec;
        }
        for (int t = tid; t < SP * nV; t += NT) {
            const int pt = t / nV, i = t - pt * nV;
            const int j0 = pt * CH, j1 = (j0 + CH < nV) ? j0 + CH : nV;
            double acc = 0.0;
            for (int j = j0; j < j1; ++j) acc += ((j <= i) ? S.Hp[G::hp(i, j)] : S.Hp[G::hp(j, i)]) * v[j];
            S.wpart[pt][i] = acc;
        }
        __syncthreads();
        for (int i = tid; i < nV; i += NT) {
            double acc = S.wpart[0][i];
#pragma unroll
            for (int pt = 1; pt < SP; ++pt) acc += S.wpart[pt][i];
            if constexpr (SM::WSP > 0) {
                if (i < SM::WSP) {
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
        __syncthreads();
    }

    // x_c = -J2 J2' g for the curved variables (flat ones keep the bound the caller put in x); barrier passed on return
    __device__ static void initial_point(SM& S, const Tile& m, int q, int nC, int nV) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        double y[CPW], ymine, z[RS];
        y_dense(S, m, S.g, y, ymine);
        z_part(S, m, y, ymine, q);
        __syncthreads();
        z_sum(S, q, z);
        if (warp == 0) {
#pragma unroll
            for (int s = 0; s < RS; ++s) {
                const int i = lane + 32 * s;
                if (i < nC) S.x[i] = -z[s];
            }
        }
        __syncthreads();
    }

    template <class Prob>
    __device__ static GiStats solve(const Prob& prob, SM& S, Tile& m, int& q, int nV, double tol, int max_iter) {
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, col0 = warp * CPW;
        const int jmine = col0 + lane;                // the column lane c < CPW looks after
        const bool colane = lane < CPW;
        GiStats st = {0, GI_EXIT_SOLVED, 0, 0, 0};
        int rbuf = 0, drops_at_refresh = 0;
        bool refresh_failed = false;
        PHASE_DECL;
        publish_srow(S, m);
        __syncwarp();

        // drop working-set column l: K1 <- K1 + k r'^T with r' = -K1' H k / k'Hk (k = column l), the freed direction
        // k / sqrt(k'Hk) joins J2 as column q-1, column q-1 of K1 moves into slot l.  Barrier passed on return.
        auto drop_column = [&](int l) {
            const int wl = l / CPW, cl = l - wl * CPW;
            if (warp == wl) {
                col_store(m, cl, S.colk[1]);
                __syncwarp();
                const double sc = S.cs[l];
#pragma unroll
                for (int s = 0; s < RS; ++s) S.colk[1][lane + 32 * s] *= sc;     // own entries: k = cs_l * (tile column l)
            }
            __syncthreads();                                  // k = M[:, l] visible block-wide
            symv(S, S.colk[1], nullptr, nV);                  // rowv = H k
            double ks[RS], kw = 0.0;
#pragma unroll
            for (int s = 0; s < RS; ++s) {
                ks[s] = S.colk[1][lane + 32 * s];             // the TRUE column (scaled at the store)
                kw = fma(ks[s], S.rowv[lane + 32 * s], kw);
            }
            const double kHk = warp_sum_d(kw);
            double rp[CPW], rpm;
            y_dense(S, m, S.rowv, rp, rpm);                   // rp_j = M[:, j]' (H k)
            const double ik = 1.0 / kHk;
            const double rs = rsqrt(kHk);
            const int q1 = q - 1, wq1 = q1 / CPW, cq1 = q1 - wq1 * CPW;
            if (col0 < q) {                                   // this warp has K1 columns
                if (colane) S.ysm[warp][1][lane] = (jmine < q && jmine != l) ? rpm * ik : 0.0;
                if (lane >= CPW && lane < CPW2) S.ysm[warp][1][lane] = 0.0;
                __syncwarp();
                double rn[CPW];
                load_y(S, 1, rn);
#pragma unroll
                for (int c = 0; c < CPW; ++c) {
#pragma unroll
                    for (int s = 0; s < RS; ++s) m(s, c) = fma(-ks[s], rn[c], m(s, c));
                }
            }
            if (warp == wq1 && l != q1) col_store(m, cq1, S.colk[0]);     // the (updated) last K1 column moves into slot l
            __syncthreads();
            if (warp == wl && l != q1) {
                double cv[RS];
#pragma unroll
                for (int s = 0; s < RS; ++s) cv[s] = S.colk[0][lane + 32 * s];
                col_set(m, cl, cv);
            }
            if (warp == wq1) {
                double cv[RS];
#pragma unroll
                for (int s = 0; s < RS; ++s) cv[s] = ks[s] * rs;
                col_set(m, cq1, cv);
            }
            publish_srow(S, m);
            if (tid == 0) {
                S.status[S.act[l] >> 1] = 0;
                S.act[l] = S.act[q1];
                S.lam[l] = S.lam[q1];
                S.lam[q1] = 0.0;
                S.cs[l] = S.cs[q1];                           // the moved column keeps its scale; the new J2 column is unscaled
                S.cs[q1] = 1.0;
            }
            --q;
            ++st.n_drop;
            __syncthreads();
        };

        while (true) {
            // P1: the most violated inactive constraint side (policy evaluates its slots; one candidate per thread)
            double best = 0.0;
            int best_i = 0x7fffffff;
            PHASE(0);
            PHASE_COUNT(12);
            prob.search(best, best_i);
            PHASE(1);
            int code;
            double viol;
            {
                unsigned long long km;
                const int wl = warp_argmin_key(dkey(best), km);
                const int wi = __shfl_sync(0xffffffffu, best_i, wl);
                if (lane == 0) { S.red_key[rbuf][warp] = km; S.red_idx[rbuf][warp] = wi; }
                __syncthreads();
                unsigned long long ck = DKEY_NONE;
                int ci = 0x7fffffff;
                if (lane < NW) { ck = S.red_key[rbuf][lane]; ci = S.red_idx[rbuf][lane]; }
                rbuf ^= 1;
                const int wl2 = warp_argmin_key(ck, km);
                code = __shfl_sync(0xffffffffu, ci, wl2);
                viol = dkey_inv(km);
            }
            PHASE(2);

            if (!(viol < -tol)) {
                // the refresh repairs what chains of partial steps leave behind; a run of pure full steps keeps x the
                // exact working-set minimiser (to round-off)
                if (st.n_drop == drops_at_refresh) break;
                ++st.n_refresh;
                drops_at_refresh = st.n_drop;
                // Newton step on the active manifold + multipliers from stationarity; a column whose recomputed
                // multiplier is negative is dropped and the step repeated (see gi_core.cuh)
                for (int pass = 0;; ++pass) {
                    symv(S, S.x, S.g, nV);                               // rowv = H x + g
                    double y[CPW], ymine, z[RS];
                    y_dense(S, m, S.rowv, y, ymine);
                    z_part(S, m, y, ymine, q);
                    double ymn = 0.0, ymx = 0.0;
                    if (colane && jmine < q) {
                        const double yt = ymine * S.cs[jmine];          // true multiplier of column jmine
                        S.lam[jmine] = fmax(yt, 0.0);
                        ymx = fabs(yt);
                        ymn = fmin(yt, 0.0);
                    }
                    {
                        unsigned long long km;
                        const int wl = warp_argmin_key(dkey(ymn), km);
                        const int lm = __shfl_sync(0xffffffffu, jmine, wl);
                        if (lane == 0) { S.red_ymin[warp] = km; S.red_lmin[warp] = lm; }
                        (void)warp_argmin_key(dkey(-ymx), km);
                        if (lane == 0) S.red_ymax[warp] = km;
                    }
                    __syncthreads();
                    z_sum(S, q, z);
                    if (warp == 0) {
#pragma unroll
                        for (int s = 0; s < RS; ++s) {
                            const int i = lane + 32 * s;
                            if (i < nV) S.x[i] -= z[s];
                        }
                    }
                    double ymin, ymax;
                    int lmin;
                    {
                        unsigned long long km, ck = (lane < NW) ? S.red_ymin[lane] : DKEY_NONE;
                        const int li = (lane < NW) ? S.red_lmin[lane] : 0;
                        const int wl = warp_argmin_key(ck, km);
                        lmin = __shfl_sync(0xffffffffu, li, wl);
                        ymin = dkey_inv(km);
                        ck = (lane < NW) ? S.red_ymax[lane] : DKEY_NONE;
                        (void)warp_argmin_key(ck, km);
                        ymax = -dkey_inv(km);
                    }
                    __syncthreads();                   // x complete
                    if (!(ymin < -1e-10 * (1.0 + ymax))) break;
                    if (pass >= 8 || ++st.iters > max_iter) { st.exitflag = GI_EXIT_MAXITER; refresh_failed = true; break; }
                    drop_column(lmin);
                    drops_at_refresh = st.n_drop;
                }
                if (refresh_failed) break;
                continue;
            }

            const int pslot = code >> 1, pside = (code & 1) ? +1 : -1;
            double sp = viol;                       // n'x - b  (< 0)
            double lam_p = 0.0;
            const double nn = prob.norm2(pslot);
            const SpN sn = prob.sparse_normal(pslot, pside);
            const int scnt = sn.cnt;
            const Sparse spn = sparse_split(sn);
            if (scnt == 0) {                        // dense normal: every warp needs all of it
                const auto prep = prob.normal_prepare(pslot, pside);
                for (int i = tid; i < G::VL; i += NT) S.nvec[i] = (i < nV) ? prob.normal_entry(prep, i) : 0.0;
                __syncthreads();
            }
            PHASE(3);
            bool failed = false;
            while (true) {
                if (++st.iters > max_iter) { st.exitflag = GI_EXIT_MAXITER; failed = true; break; }
                // P3: y = M'n for this warp's columns
                double y[CPW], ymine;
                if (scnt > 0) y_sparse(S, m, spn, y, ymine);
                else y_dense(S, m, S.nvec, y, ymine);
                // partial step-length quantities of the warp's columns (lane c: column col0 + c)
                const int wq = q / CPW, cq = q - wq * CPW;
                double ytrue;
                {
                    const bool isK = colane & (jmine < q);
                    const bool isJ = colane & (jmine >= q) & (jmine < nV);
                    ytrue = isK ? ymine * S.cs[isK ? jmine : 0] : ymine;      // K1 columns carry a scale (J2 columns: 1)
                    const bool cand = isK & (ytrue > 1e-13);
                    const double d2p = warp_sum_d(isJ ? ymine * ymine : 0.0);
                    const double tj = cand ? S.lam[cand ? jmine : 0] * __drcp_rn(cand ? ytrue : 1.0) : INFINITY;
                    unsigned long long km;
                    const int wl = warp_argmin_key(dkey(tj), km);
                    const int lj = __shfl_sync(0xffffffffu, jmine, wl);
                    if (lane == 0) { S.red_d2[warp] = d2p; S.red_t1[warp] = km; S.red_l[warp] = lj; }
                    if (warp == wq && lane == cq) S.red_yq = ymine;
                }
                // z = J2 y2 (partial over the warp's columns); the column that leaves J2 (column q)
                z_part(S, m, y, ymine, q);
                if (warp == wq) col_store(m, cq, S.colk[0]);
                PHASE(5);
                __syncthreads();
                double z[RS];
                z_sum(S, q, z);
                double d2 = 0.0, t1;
                int l;
                {
#pragma unroll
                    for (int w = 0; w < NW; ++w) d2 += S.red_d2[w];
                    unsigned long long km, ck = (lane < NW) ? S.red_t1[lane] : dkey(INFINITY);
                    const int li = (lane < NW) ? S.red_l[lane] : -1;
                    const int wl = warp_argmin_key(ck, km);
                    l = __shfl_sync(0xffffffffu, li, wl);
                    t1 = dkey_inv(km);
                }
                const bool lin_dep = !(d2 > 1e-13 * fmax(1.0, nn));
                const double rs2 = rsqrt(d2);
                const double inv_d2 = rs2 * rs2;            // (2-3 ulp: only scales step lengths and the new column)
                const double yqv = S.red_yq;
                const double delta = d2 * rs2;
                const double sgd = (yqv >= 0.0) ? delta : -delta;
                const double beta = __drcp_rn(d2 + fabs(yqv) * delta);
                const double t2 = lin_dep ? INFINITY : (sp < 0.0 ? -sp * inv_d2 : 0.0);
                const bool full = (t2 <= t1);
                if (isinf(t1) && isinf(t2)) { st.exitflag = GI_EXIT_INFEASIBLE; failed = true; break; }
                const bool primal = !isinf(t2);
                const double t = full ? t2 : t1;
                PHASE(6);
                // x += t z (warp 0), lam -= t y (every warp: its columns)
                if (primal) {
                    if (warp == 0) {
#pragma unroll
                        for (int s = 0; s < RS; ++s) {
                            const int i = lane + 32 * s;
                            if (i < nV) S.x[i] = fma(t, z[s], S.x[i]);
                        }
                    }
                    sp += t * d2;
                }
                if (colane && jmine < q) S.lam[jmine] -= t * ytrue;
                lam_p += t;
                if (full && tid == 0) {
                    S.act[q] = pslot * 2 + (pside > 0 ? 1 : 0);
                    S.status[pslot] = (int8_t)pside;
                    S.lam[q] = lam_p;
                    S.cs[q] = -sgd * inv_d2;       // column q becomes the Householder image -z / sgd = k / cs[q],  k = z / d2
                }
                PHASE(7);
                if (full) {
                    // add p:  K1 <- K1 - k r',  J2 <- J2 (I - beta v v'),  column q <- k = z / d2
                    double kr[RS], wr[RS];
#pragma unroll
                    for (int s = 0; s < RS; ++s) {
                        kr[s] = z[s] * inv_d2;
                        wr[s] = (z[s] + sgd * S.colk[0][lane + 32 * s]) * beta;
                    }
                    if (warp < wq) {                    // all columns left of q:  m - k r'
#pragma unroll
                        for (int c = 0; c < CPW; ++c) {
#pragma unroll
                            for (int s = 0; s < RS; ++s) m(s, c) = fma(-kr[s], y[c], m(s, c));
                        }
                    } else if (warp > wq) {             // all columns right of q:  m - w y'
#pragma unroll
                        for (int c = 0; c < CPW; ++c) {
#pragma unroll
                            for (int s = 0; s < RS; ++s) m(s, c) = fma(-wr[s], y[c], m(s, c));
                        }
                    } else {
                        // the warp that holds column q:  m - k yK' - w yJ'  with yK = y left of q, yJ = the Householder
                        // vector right of q and y_q + sgd AT q.  Column q thereby becomes the image -z / sgd of the
                        // reflection, which is k up to the scale recorded in cs[q]: no register is written at a
                        // run-time index.  The two masked vectors come from the column lanes through shared memory.
                        if (lane < CPW2) {
                            const bool in = lane < CPW;
                            S.ysm[warp][1][lane] = (in && lane < cq) ? ymine : 0.0;
                            S.ysm[warp][2][lane] = (in && lane > cq) ? ymine : ((in && lane == cq) ? ymine + sgd : 0.0);
                        }
                        __syncwarp();
                        double ya[CPW], yb[CPW];
                        load_y(S, 1, ya);
                        load_y(S, 2, yb);
#pragma unroll
                        for (int c = 0; c < CPW; ++c) {
#pragma unroll
                            for (int s = 0; s < RS; ++s) m(s, c) = fma(-wr[s], yb[c], fma(-kr[s], ya[c], m(s, c)));
                        }
                    }
                    publish_srow(S, m);
                    ++q;
                    ++st.n_add;
                    PHASE(8);
                    break;
                }
                // partial step: drop the blocking constraint l and go on with the same p
                drop_column(l);                        // (its barriers also publish lam and x of this step)
                PHASE(10);
            }
            if (failed) break;
            __syncthreads();                           // x, act, status, lam of the step published
        }
        __syncthreads();
        PHASE_FLUSH;
        return st;
    }

    // 1/2 x'Hx + g'x with the packed H; block-uniform result
    __device__ static double objective(SM& S, int nV) {
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        symv(S, S.x, nullptr, nV);
        double acc = 0.0;
        for (int i = tid; i < nV; i += NT) acc += S.x[i] * (0.5 * S.rowv[i] + S.g[i]);
        acc = warp_sum_d(acc);
        if (lane == 0) S.red_val[warp] = acc;
        __syncthreads();
        double f = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) f += S.red_val[w];
        __syncthreads();
        return f;
    }
};

}  // namespace fsae
