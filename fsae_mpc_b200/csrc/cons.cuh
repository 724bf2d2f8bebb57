// Constraint policies: how the reference's *_state_constraints.m rows are represented
// on chip without ever materialising the dense xA matrix.
//
// Every general constraint of the reference is a "row" r at horizon step k:
//       lo <= rowval_r,k(u) (+ slack)          [lower side, qpOASES status -1]
//             rowval_r,k(u) (- slack) <= up    [upper side, qpOASES status +1]
// where rowval is a linear form of the predicted state at step k (and, for the dynamic
// friction polygon, of u_k).  The reference writes the two sides of a soft row as two
// separate xA rows with a +-1e10 far bound on the other side
// (kinematic_state_constraints.m:38-48, dynamic_state_constraints.m:38-57); those far
// sides can never be active, so one two-sided row with a side-dependent slack sign is the
// same feasible set and the same multipliers.
//
// Predicted-state perturbations are xs[c][k] = (B_bar u)[state c, step k] for the states
// the constraints touch.  "Real" states need their B_bar rows (kept packed in shared
// memory); "integrator" states (continuous dynamics x' = u_j exactly: v, delta) have
// B_bar rows that are exactly dt on the controls up to step k, so they are prefix sums.
#pragma once
#include "models.cuh"
#include "gi_core.cuh"   // SpN

namespace fsae {

template <class Model>
struct Cons;

// ---------------------------------------------------------------- kinematic
// rows per step: 0 v (hard, lower only in the reference: ub = inf), 1 delta (hard),
//                2 n (soft, slack 0), 3 tyre ay (soft, slack 0)
template <>
struct Cons<KinModel> {
    static constexpr int NR = 4;          // rows per step
    static constexpr int NREAL = 3;       // states whose B_bar rows are built (s, n, mu)
    static constexpr int NCR = 1;         // of those, rows the constraints need (n)
    static constexpr int NINT = 2;        // integrator states (v <- u0, delta <- u1)
    static constexpr int NPC = 2;         // per-step coefficients (tyre: c4, c5)
    static constexpr int NG0 = 1;         // per-step constraint offsets (tyre g0)
    static constexpr int NCG = 1;         // per-problem constants (none for this model)
    __device__ static void problem_consts(const fsae_params& p, double* cg) { cg[0] = 0.0; }
    static constexpr int NXS = NCR + NINT;
    // Every control is integrated exactly by one state (v' = u0, delta' = u1): the solver works in the
    // integrator coordinates w = T u (w_{2k+c} = dt * sum_{i<=k} u_{2i+c} = the perturbation of v / delta at
    // step k), where all rows but the n rows have at most three entries (fused_v2.cuh).
    // (The product kernels work in control coordinates: WSPACE is switched on by the model tags KinModelW / KinModelR
    // below, which the cross-check library instantiates.  Measured on B200, DESIGN.md "Integrator coordinates": the
    // leaner formulation is 3-7 % SLOWER than control coordinates at horizons 20 and 40.)
    static constexpr bool WSPACE = false;
    static constexpr bool USE_RL = false;  // row-lane core (gi_core_rl.cuh); needs WSPACE
    // w-space normal of row r at step k in  n'x >= b  form (sg = +1 lower side, -1 upper side): entries in
    // ascending order, 0 = dense row.  nU = index of slack 0.
    __device__ __forceinline__ static SpN sparse_row(int r, int k, const double* pc, double sg, int nU) {
        switch (r) {
            case 0: return SpN{1, 2 * k, 0, 0, sg, 0.0, 0.0};
            case 1: return SpN{1, 2 * k + 1, 0, 0, sg, 0.0, 0.0};
            case 2: return SpN{0, 0, 0, 0, 0.0, 0.0, 0.0};
            default: return SpN{3, 2 * k, 2 * k + 1, nU, sg * pc[0], sg * pc[1], 1.0};
        }
    }
    __host__ __device__ static constexpr int real_state(int i) { return i; }        // 0,1,2
    __host__ __device__ static constexpr int cons_real(int i) { return 1; }         // n is real row 1
    __host__ __device__ static constexpr int int_state(int i) { return 3 + i; }     // 3,4
    __host__ __device__ static constexpr int int_ucol(int i) { return i; }
    // number of reference xA rows: 6*N
    __host__ __device__ static constexpr int n_ref_rows(int N) { return 6 * N; }

    // slack variable index (0-based within the slack block) used by row r, or -1
    __device__ __forceinline__ static int row_slack(int r) { return r >= 2 ? 0 : -1; }

    // per-step coefficients from the linearisation point
    // (kinematic_tyre_linearise_constraints.m:18-26): C = [0 0 0 2 v d, v^2]/(lf+lr), g0 = v^2 d/(lr+lf)
    __device__ static void step_coefs(const double* xl, const double* ul, const DevTrack& tr,
                                      const fsae_params& p, double* pc, double* g0, const KinModel::Aux* = nullptr) {
        const double L = p.lr + p.lf;
        pc[0] = 2.0 * xl[3] * xl[4] / L;
        pc[1] = xl[3] * xl[3] / L;
        g0[0] = xl[3] * xl[3] * xl[4] / L;
    }

    // row value from xs = [n, v, delta] perturbations at step k
    __device__ __forceinline__ static double row_value(int r, const double* xs, const double* pc,
                                                       const double* cg, double ua) {
        switch (r) {
            case 0: return xs[1];
            case 1: return xs[2];
            case 2: return xs[0];
            default: return pc[0] * xs[1] + pc[1] * xs[2];
        }
    }
    // coefficient of xs[c] in row r (c: 0 n, 1 v, 2 delta) and of u_a,k
    __device__ __forceinline__ static double row_coef(int r, int c, const double* pc, const double* cg) {
        switch (r) {
            case 0: return c == 1 ? 1.0 : 0.0;
            case 1: return c == 2 ? 1.0 : 0.0;
            case 2: return c == 0 ? 1.0 : 0.0;
            default: return c == 1 ? pc[0] : (c == 2 ? pc[1] : 0.0);
        }
    }
    __device__ __forceinline__ static double row_ucoef(int r, int uc, const double* pc, const double* cg) { return 0.0; }

    // bounds of row r at step k given the free response xf (= A_bar x0 + d_bar at step k),
    // the linearisation point and g0  (kinematic_state_constraints.m:29-39,
    // kinematic_tyre_linearise_constraints.m:30-32)
    __device__ static void row_bounds(int r, const double* xf, const double* xl, const double* ul,
                                      const double* pc, const double* g0, const double* cg,
                                      const fsae_params& p, double& lo, double& up) {
        switch (r) {
            case 0: lo = p.vel_lb - xf[3]; up = p.vel_ub - xf[3]; break;
            case 1: lo = p.delta_lb - xf[4]; up = p.delta_ub - xf[4]; break;
            case 2: lo = p.n_lb - xf[1]; up = p.n_ub - xf[1]; break;
            default: {
                const double c = g0[0] + pc[0] * (xf[3] - xl[3]) + pc[1] * (xf[4] - xl[4]);
                lo = -p.ay_max - c;
                up = p.ay_max - c;
            }
        }
    }
    // full-state index of constraint state c (xs ordering: n, v, delta)
    __host__ __device__ static constexpr int xs_state(int c) { return c == 0 ? 1 : (c == 1 ? 3 : 4); }
    // Decode a reference xA row index into (row r, step k, kind).  kind: 0 hard two-sided,
    // 1 lower side + slack with far upper, 2 upper side - slack with far lower,
    // 3 lower side + slack with +inf upper, 4 upper side - slack with -inf lower.
    __device__ __forceinline__ static void ref_decode(int row, int N, int& r, int& k, int& kind) {
        const int blk = row / N;
        k = row - blk * N;
        switch (blk) {
            case 0: r = 0; kind = 0; break;
            case 1: r = 1; kind = 0; break;
            case 2: r = 2; kind = 1; break;
            case 3: r = 2; kind = 2; break;
            case 4: r = 3; kind = 3; break;
            default: r = 3; kind = 4; break;
        }
    }
    // slack-column entry of a reference xA row (kinematic_state_constraints.m:42,48)
    __device__ __forceinline__ static double ref_slack_sign(int row, int N) {
        const int blk = row / N;
        return (blk == 2 || blk == 4) ? 1.0 : -1.0;
    }
    // index of (row r, step k, side) in the reference's xA row order
    // [v | delta | n+s | n-s | ay+s | ay-s], each block N long.
    __device__ __forceinline__ static int ref_row(int r, int k, int side, int N) {
        switch (r) {
            case 0: return k;
            case 1: return N + k;
            case 2: return (side < 0 ? 2 * N : 3 * N) + k;
            default: return (side < 0 ? 4 * N : 5 * N) + k;
        }
    }
};


// The kinematic model in integrator coordinates: with the column-lane core (KinModelW) and with the row-lane core
// (KinModelR).  Same model, same constraints; only the solver's coordinates and operator layout differ.
struct KinModelW : KinModel {};
struct KinModelR : KinModel {};
template <>
struct Cons<KinModelW> : Cons<KinModel> {
    static constexpr bool WSPACE = true;
};
template <>
struct Cons<KinModelR> : Cons<KinModel> {
    static constexpr bool WSPACE = true;
    static constexpr bool USE_RL = true;
};

// ---------------------------------------------------------------- dynamic
// rows per step: 0 x_d (hard), 1 delta (hard), 2 n (soft, slack 0), 3 alpha_r (soft, slack 1),
//                4 alpha_f (soft, slack 2), 5..16 friction-polygon edges (upper only, slack 3)
// (dynamic_state_constraints.m:1-58, dynamic_slip_linearise_constraints.m:1-47,
//  dynamic_tyre_linearise_constraints.m:1-64)
template <>
struct Cons<DynModel> {
    static constexpr int NPOLY = 12;      // dynamic_tyre_linearise_constraints.m:18
    static constexpr int NR = 5 + NPOLY;
    static constexpr int NREAL = 6;       // s, n, mu, x_d, y_d, theta_d
    static constexpr int NCR = 4;         // constraints touch n, x_d, y_d, theta_d
    static constexpr int NINT = 1;        // delta <- u1
    static constexpr int NPC = 9;         // alpha_r (3), alpha_f (3), tyre base (3) on (x_d, y_d, theta_d)
    static constexpr int NG0 = 3;         // -atan(vr), delta - atan(vf), Fcr/m
    static constexpr int NCG = 4 * NPOLY; // ac_list, al_list, dac, dal
    static constexpr int NXS = NCR + NINT;
    static constexpr bool WSPACE = false;      // only delta is an integrator state; the rows that cycle are dense
    static constexpr bool USE_RL = false;
    __device__ __forceinline__ static SpN sparse_row(int, int, const double*, double, int) { return SpN{0, 0, 0, 0, 0.0, 0.0, 0.0}; }
    __host__ __device__ static constexpr int real_state(int i) { return i; }
    __host__ __device__ static constexpr int cons_real(int i) { return i == 0 ? 1 : i + 2; }   // 1,3,4,5
    __host__ __device__ static constexpr int int_state(int i) { return 6; }
    __host__ __device__ static constexpr int int_ucol(int i) { return 1; }
    __host__ __device__ static constexpr int n_ref_rows(int N) { return 20 * N; }
    // xs ordering: 0 n, 1 x_d, 2 y_d, 3 theta_d, 4 delta
    __host__ __device__ static constexpr int xs_state(int c) { return c == 0 ? 1 : (c + 2); }

    __device__ __forceinline__ static int row_slack(int r) { return r < 2 ? -1 : (r < 5 ? r - 2 : 3); }

    // polygon vertices/edges (dynamic_tyre_linearise_constraints.m:18-23)
    __device__ static void problem_consts(const fsae_params& p, double* cg) {
        for (int j = 0; j < NPOLY; ++j) {
            const double t0 = (2.0 * M_PI) * (double)j / NPOLY, t1 = (2.0 * M_PI) * (double)(j + 1) / NPOLY;
            const double ac0 = p.ac_max * sin(t0), ac1 = p.ac_max * sin(t1);
            const double al0 = p.al_max * cos(t0), al1 = p.al_max * cos(t1);
            cg[j] = ac0;
            cg[NPOLY + j] = al0;
            cg[2 * NPOLY + j] = ac1 - ac0;
            cg[3 * NPOLY + j] = al1 - al0;
        }
    }

    // aux: the tyre-force terms of the model evaluation AT (xl, ul), if the caller has them (linearise_step hands out
    // those of its first stage: one model evaluation per step saved); else they are evaluated here
    __device__ static void step_coefs(const double* xl, const double* ul, const DevTrack& tr,
                                      const fsae_params& p, double* pc, double* g0, const DynAux* aux = nullptr) {
        DynAux a;
        if (aux) a = *aux;
        else DynModel::eval_aux(xl, ul, tr, p, nullptr, nullptr, &a);
        const double ih = 1.0 / a.x_d_hat;
        pc[0] = a.denom_vr2 * a.vr * a.x_d_hat_d * ih;
        pc[1] = -a.denom_vr2 * ih;
        pc[2] = a.denom_vr2 * p.lr * ih;
        pc[3] = a.denom_vf2 * a.vf * a.x_d_hat_d * ih;
        pc[4] = -a.denom_vf2 * ih;
        pc[5] = -a.denom_vf2 * p.lf * ih;
        pc[6] = -a.Fcr_d * a.denom_vr2 * a.vr * a.x_d_hat_d * ih / p.mass;
        pc[7] = a.Fcr_d * a.denom_vr2 * ih / p.mass;
        pc[8] = -a.Fcr_d * a.denom_vr2 * p.lr * ih / p.mass;
        g0[0] = -atan(a.vr);
        g0[1] = xl[6] - atan(a.vf);
        g0[2] = a.Fcr / p.mass;
    }

    __device__ __forceinline__ static double row_value(int r, const double* xs, const double* pc,
                                                       const double* cg, double ua) {
        switch (r) {
            case 0: return xs[1];
            case 1: return xs[4];
            case 2: return xs[0];
            case 3: return pc[0] * xs[1] + pc[1] * xs[2] + pc[2] * xs[3];
            case 4: return pc[3] * xs[1] + pc[4] * xs[2] + pc[5] * xs[3] + xs[4];
            default: {
                const int j = r - 5;
                return cg[3 * NPOLY + j] * (pc[6] * xs[1] + pc[7] * xs[2] + pc[8] * xs[3]) + cg[2 * NPOLY + j] * ua;
            }
        }
    }
    __device__ __forceinline__ static double row_coef(int r, int c, const double* pc, const double* cg) {
        switch (r) {
            case 0: return c == 1 ? 1.0 : 0.0;
            case 1: return c == 4 ? 1.0 : 0.0;
            case 2: return c == 0 ? 1.0 : 0.0;
            case 3: return (c >= 1 && c <= 3) ? pc[c - 1] : 0.0;
            case 4: return (c >= 1 && c <= 3) ? pc[3 + c - 1] : (c == 4 ? 1.0 : 0.0);
            default: return (c >= 1 && c <= 3) ? cg[3 * NPOLY + (r - 5)] * pc[6 + c - 1] : 0.0;
        }
    }
    __device__ __forceinline__ static double row_ucoef(int r, int uc, const double* pc, const double* cg) {
        return (r >= 5 && uc == 0) ? cg[2 * NPOLY + (r - 5)] : 0.0;
    }
    __device__ static void row_bounds(int r, const double* xf, const double* xl, const double* ul,
                                      const double* pc, const double* g0, const double* cg,
                                      const fsae_params& p, double& lo, double& up) {
        const double d3 = xf[3] - xl[3], d4 = xf[4] - xl[4], d5 = xf[5] - xl[5];
        switch (r) {
            case 0: lo = p.vel_lb - xf[3]; up = p.vel_ub - xf[3]; break;
            case 1: lo = p.delta_lb - xf[6]; up = p.delta_ub - xf[6]; break;
            case 2: lo = p.n_lb - xf[1]; up = p.n_ub - xf[1]; break;
            case 3: {
                const double c = g0[0] + pc[0] * d3 + pc[1] * d4 + pc[2] * d5;
                lo = -p.slip_max - c; up = p.slip_max - c; break;
            }
            case 4: {
                const double c = g0[1] + pc[3] * d3 + pc[4] * d4 + pc[5] * d5 + (xf[6] - xl[6]);
                lo = -p.slip_max - c; up = p.slip_max - c; break;
            }
            default: {
                const int j = r - 5;
                const double dac = cg[2 * NPOLY + j], dal = cg[3 * NPOLY + j];
                const double gj = (ul[0] - cg[NPOLY + j]) * dac - (g0[2] - cg[j]) * dal;
                const double c = gj + dal * (pc[6] * d3 + pc[7] * d4 + pc[8] * d5) - dac * ul[0];
                lo = -INFINITY; up = 0.0 - c;
            }
        }
    }
    __device__ __forceinline__ static void ref_decode(int row, int N, int& r, int& k, int& kind) {
        if (row < N) { r = 0; k = row; kind = 0; }
        else if (row < 2 * N) { r = 1; k = row - N; kind = 0; }
        else if (row < 3 * N) { r = 2; k = row - 2 * N; kind = 1; }
        else if (row < 4 * N) { r = 2; k = row - 3 * N; kind = 2; }
        else if (row < 6 * N) { const int t = row - 4 * N; k = t >> 1; r = 3 + (t & 1); kind = 3; }
        else if (row < 8 * N) { const int t = row - 6 * N; k = t >> 1; r = 3 + (t & 1); kind = 4; }
        else { const int t = row - 8 * N; k = t / NPOLY; r = 5 + (t - k * NPOLY); kind = 4; }
    }
    __device__ __forceinline__ static double ref_slack_sign(int row, int N) {
        int r, k, kind;
        ref_decode(row, N, r, k, kind);
        return (kind == 1 || kind == 3) ? 1.0 : -1.0;
    }
    __device__ __forceinline__ static int ref_row(int r, int k, int side, int N) {
        switch (r) {
            case 0: return k;
            case 1: return N + k;
            case 2: return (side < 0 ? 2 * N : 3 * N) + k;
            case 3: return (side < 0 ? 4 * N : 6 * N) + 2 * k;
            case 4: return (side < 0 ? 4 * N : 6 * N) + 2 * k + 1;
            default: return 8 * N + NPOLY * k + (r - 5);
        }
    }
};

}  // namespace fsae
