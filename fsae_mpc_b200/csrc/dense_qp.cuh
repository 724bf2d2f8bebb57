// Batched dense QP solve: drop-in for
//     [x,fval,exitflag,iter,lambda,auxOutput] = qpOASES(H,g,A,lb,ub,lbA,ubA)
// (optimizers/matlab/qpOASES/qpOASES.m:22) for B independent QPs of one shape.  Two instantiations: nV <= 95
// (8 warps, operator in registers, packed H in shared memory) and nV <= 191 (12 warps, half of the operator tile in
// shared memory, H as a full symmetric matrix in a per-problem global slab that stays L2-resident) -- the second
// covers the condensed QPs of horizon 80 (nV = 161 / 164).
// Same register-tiled dual active-set core as the fused MPC kernel (gi_core.cuh); here the
// problem policy reads the dense constraint matrix from global memory (column-major
// [nC x nV], so consecutive threads read consecutive rows: coalesced).
//
// Variables whose Hessian row is zero (pure linear cost, e.g. exact-penalty slacks) are
// ordered last internally and start on the bound their gradient pushes them to; this is what
// keeps 1e8-size penalty gradients out of the iterate.
#pragma once
#include "gi_core.cuh"
#include "../../include/fsae_mpc_b200.h"

namespace fsae {

struct DenseArgs {
    int B, nV, nC;
    const double *H, *g, *A, *lb, *ub, *lbA, *ubA;
    double *x, *fval, *lambda;
    int32_t *exitflag, *iters;
    int8_t *wsB, *wsC;
    double feas_tol, flat_eps;
    int max_iter;
    unsigned long long* counters;
    double* hscratch;       // large instantiation: B x nV x nV doubles (H in the solver's variable order)
};

template <int NVMAX, int NW = 8, int CSR = -1, bool HPG = false>
struct DenseSm {
    using G = GiCfg<NVMAX, NW, 1, CSR>;
    double lbv[NVMAX], ubv[NVMAX];     // variable bounds in internal order
    int perm[NVMAX];                   // internal index -> caller's index
    int nflat, ncurv;
    int pad_[2];
    alignas(16) double Msm[GiTile<G>::SM_DOUBLES > 0 ? GiTile<G>::SM_DOUBLES : 2];   // shared part of the operator tiles
    GiSm<G, 8, HPG> gi;                // status[] continues into the dynamic tail (nV + nC bytes)
};

template <int NVMAX, int NW = 8, int CSR = -1, bool HPG = false>
struct DenseProb {
    using G = GiCfg<NVMAX, NW, 1, CSR>;
    static constexpr bool REUSE = false;       // no candidate reuse: a row evaluation is a full global-memory dot product
    __device__ __forceinline__ double eval_code(int) const { return 0.0; }
    DenseSm<NVMAX, NW, CSR, HPG>& S;
    const double* A;
    const double* lbA;
    const double* ubA;
    int nV, nC;

    __device__ __forceinline__ void search(double& best, int& best_i) const {
        const int tid = threadIdx.x;
        const double* x = S.gi.x;
        for (int slot = tid; slot < nV; slot += G::NT) {
            if (S.gi.status[slot] != 0) continue;
            const double vlo = x[slot] - S.lbv[slot], vup = S.ubv[slot] - x[slot];
            if (vlo < best) { best = vlo; best_i = slot * 2; }
            if (vup < best) { best = vup; best_i = slot * 2 + 1; }
        }
        for (int r = tid; r < nC; r += G::NT) {
            const int slot = nV + r;
            if (S.gi.status[slot] != 0) continue;
            double acc = 0.0;
            for (int i = 0; i < nV; ++i) acc = fma(A[(size_t)S.perm[i] * nC + r], x[i], acc);
            const double vlo = acc - lbA[r], vup = ubA[r] - acc;
            if (vlo < best) { best = vlo; best_i = slot * 2; }
            if (vup < best) { best = vup; best_i = slot * 2 + 1; }
        }
    }
    struct Prep { int pslot; double sg; };
    __device__ __forceinline__ Prep normal_prepare(int pslot, int pside) const { return {pslot, pside < 0 ? 1.0 : -1.0}; }
    __device__ __forceinline__ double normal_entry(const Prep& p, int i) const {
        if (p.pslot < nV) return (i == p.pslot) ? p.sg : 0.0;
        return p.sg * A[(size_t)S.perm[i] * nC + (p.pslot - nV)];
    }
    // variable bounds are unit normals: one entry (gi_core.cuh, sparse normals); general rows are dense
    __device__ __forceinline__ SpN sparse_normal(int pslot, int pside) const {
        if (pslot >= nV) return spn_none();
        return SpN{1, pslot, 0, 0, pside < 0 ? 1.0 : -1.0, 0.0, 0.0};
    }
    __device__ __forceinline__ double norm2(int) const { return 1.0; }
    __device__ __forceinline__ bool is_unit(int pslot) const { return pslot < nV; }
};

template <int NVMAX, int NW = 8, int CSR = -1, bool HPG = false>
__global__ void __launch_bounds__(32 * NW, 1) dense_qp_kernel(DenseArgs a) {
    using G = GiCfg<NVMAX, NW, 1, CSR>;
    using SM = GiSm<G, 8, HPG>;
    using Ops = GiOps<G, SM>;
    using DS = DenseSm<NVMAX, NW, CSR, HPG>;
    constexpr int RPW = G::RPW, CS = G::CS, NT = G::NT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DS& S = *reinterpret_cast<DS*>(smem_raw);
    SM& Q = S.gi;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, row0 = warp * RPW;
    const int b = blockIdx.x, nV = a.nV, nC = a.nC;
    const double* H = a.H + (size_t)b * nV * nV;
    const double* g = a.g + (size_t)b * nV;
    const double* A = a.A + (size_t)b * nC * nV;
    const double* lb = a.lb + (size_t)b * nV;
    const double* ub = a.ub + (size_t)b * nV;
    const double* lbA = a.lbA + (size_t)b * nC;
    const double* ubA = a.ubA + (size_t)b * nC;

    if (HPG && tid == 0) Q.hpg = a.hscratch + (size_t)b * nV * nV;
    for (int i = tid; i < nV + nC; i += NT) Q.status[i] = 0;
    for (int i = tid; i < G::RP; i += NT) { Q.x[i] = 0.0; Q.g[i] = 0.0; Q.rowv[i] = 0.0; Q.nvec[0][i] = 0.0; Q.zrow[i] = 0.0; Q.colk[0][i] = 0.0; Q.colk[1][i] = 0.0; }
    if (tid == 0) {
        // curved variables first (caller's order), zero-curvature ones last
        int nc = 0, nf = 0;
        for (int i = 0; i < nV; ++i)
            if (H[(size_t)i * nV + i] != 0.0) S.perm[nc++] = i;
        for (int i = 0; i < nV; ++i)
            if (H[(size_t)i * nV + i] == 0.0) S.perm[nc + nf++] = i;
        S.ncurv = nc;
        S.nflat = nf;
    }
    __syncthreads();
    const int nCv = S.ncurv, ns = S.nflat;
    for (int i = tid; i < nV; i += NT) {
        const int o = S.perm[i];
        Q.g[i] = g[o];
        S.lbv[i] = lb[o];
        S.ubv[i] = ub[o];
    }
    __syncthreads();
    int bad = 0;
    if (tid < ns) {
        // flat variable: sits on the bound its gradient pushes it to, multiplier |g|
        const int i = nCv + tid;
        const int side = (Q.g[i] > 0.0 || (Q.g[i] == 0.0 && isfinite(S.lbv[i]))) ? -1 : +1;
        const double bv = side < 0 ? S.lbv[i] : S.ubv[i];
        if (!isfinite(bv) || fabs(bv) >= 1e20) bad = 1;      // linear cost, no bound: unbounded below
        Q.x[i] = bv;
        Q.act[tid] = i * 2 + (side > 0 ? 1 : 0);
        Q.status[i] = (int8_t)side;
    }
    // H tiles (internal order) and the packed copy
    GiTile<G> m;
    if (G::CSS > 0) m.attach(S.Msm);
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int i = row0 + r;
#pragma unroll
        for (int s = 0; s < CS; ++s) {
            const int j = lane + 32 * s;
            double v = 0.0;
            if (i < nCv && j < nCv) v = H[(size_t)S.perm[j] * nV + S.perm[i]];
            m(r, s) = v;
            if (HPG) {          // the full symmetric matrix (every (i, j) is some thread's tile entry)
                if (i < nV && j < nV) a.hscratch[((size_t)b * nV + i) * nV + j] = (i < nCv && j < nCv) ? v : (i == j ? a.flat_eps : 0.0);
            } else {
                if (i < nV && j <= i) Q.Hp[G::hp(i, j)] = (i < nCv) ? v : (i == j ? a.flat_eps : 0.0);
            }
        }
    }
    bad = __syncthreads_or(bad);
    double lam[CS];
    int q = 0, ybuf = 0;
    const bool spd = Ops::factor_and_layout(Q, m, lam, q, nCv, ns, nV);
    GiStats st = {0, FSAE_EXIT_INTERNAL, 0, 0, 0};
    if (spd && !bad) {
        __syncthreads();
        Ops::initial_point(Q, m, ybuf, q, nCv, nV);
        const DenseProb<NVMAX, NW, CSR, HPG> prob{S, A, lbA, ubA, nV, nC};
        st = Ops::solve(prob, Q, m, lam, q, ybuf, nV, a.feas_tol, a.max_iter > 0 ? a.max_iter : 5 * (nV + nC));
    }
    __syncthreads();
    // outputs
    double f = Ops::objective(Q, nV, nullptr);
    if (tid == 0) {
        for (int i = nCv; i < nV; ++i) f -= 0.5 * a.flat_eps * Q.x[i] * Q.x[i];
        a.fval[b] = f;
        a.exitflag[b] = st.exitflag;
        if (a.iters) a.iters[b] = st.iters;
        if (a.counters) {
            atomicAdd(a.counters + 0, (unsigned long long)st.n_add);
            atomicAdd(a.counters + 1, (unsigned long long)st.n_drop);
            atomicAdd(a.counters + 2, (unsigned long long)st.n_refresh);
        }
    }
    for (int i = tid; i < nV; i += NT) a.x[(size_t)b * nV + S.perm[i]] = Q.x[i];
    if (a.lambda) {
        double* L = a.lambda + (size_t)b * (nV + nC);
        for (int i = tid; i < nV + nC; i += NT) L[i] = 0.0;
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int s = 0; s < CS; ++s) {
                const int j = lane + 32 * s;
                if (j < q) {
                    const int code = Q.act[j], slot = code >> 1;
                    const int o = slot < nV ? S.perm[slot] : slot;
                    L[o] = (code & 1) ? -lam[s] : lam[s];       // qpOASES sign: >= 0 at lower, <= 0 at upper
                }
            }
        }
    }
    if (a.wsB) {
        for (int i = tid; i < nV; i += NT) a.wsB[(size_t)b * nV + S.perm[i]] = Q.status[i];
    }
    if (a.wsC) {
        for (int r = tid; r < nC; r += NT) a.wsC[(size_t)b * nC + r] = Q.status[nV + r];
    }
}

}  // namespace fsae
