// Arguments and compile-time dimensions shared by the fused LTV-MPC kernels.
#pragma once
#include "cons.cuh"

namespace fsae {

struct BatchArgs {
    int B, N;
    double dt;
    const int32_t* track_id;
    const int32_t* param_id;
    const double* x0;
    const double* x_ref;
    const double* x_lin;
    const double* u_lin;
    double* u_opt;
    double* x_opt;
    int32_t* exitflag;
    double* fval;
    double* slack_opt;
    int32_t* iters;
    int8_t* wsB;
    int8_t* wsC;
    const DevTrack* tracks;
    const fsae_params* params;
    // optional debug taps (tests): H [nV x nV] column-major, g [nV], per problem
    double* dbg_H;
    double* dbg_g;
    double* dbg_M;                  // initial operator M = [e_slack | J] [nV x nV] column-major (register-tiled kernel)
    unsigned long long* counters;   // [0] adds, [1] drops, [2] refreshes (atomicAdd per problem)
    double* m_scratch;              // per-CTA operator slabs for the long-horizon (global-operator) variant
};

template <class Model, int N_>
struct Dims {
    using C = Cons<Model>;
    static constexpr int N = N_;
    static constexpr int NX = Model::NX, NU = Model::NU, NS = Model::NS;
    static constexpr int nU = NU * N, nV = nU + NS;
    static constexpr int LD = (nV % 2) ? nV : nV + 1;     // odd leading dim: conflict-free both ways
    static constexpr int NPK = NU * N * (N + 1) / 2;        // packed entries of one B_bar state row
    static constexpr int NROWS = C::NR * N;
    static constexpr int NSLOT = nV + NROWS;
    static constexpr int HP = nV * (nV + 1) / 2;
    __host__ __device__ static constexpr int pk(int k, int j) { return NU * k * (k + 1) / 2 + j; }
    __host__ __device__ static constexpr int hp(int i, int j) { return i * (i + 1) / 2 + j; }  // i >= j
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace fsae
