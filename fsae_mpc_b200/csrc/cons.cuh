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
    static constexpr int NXS = NCR + NINT;
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
                                      const fsae_params& p, double* pc, double* g0) {
        const double L = p.lr + p.lf;
        pc[0] = 2.0 * xl[3] * xl[4] / L;
        pc[1] = xl[3] * xl[3] / L;
        g0[0] = xl[3] * xl[3] * xl[4] / L;
    }

    // row value from xs = [n, v, delta] perturbations at step k
    __device__ __forceinline__ static double row_value(int r, const double* xs, const double* pc,
                                                       double ua) {
        switch (r) {
            case 0: return xs[1];
            case 1: return xs[2];
            case 2: return xs[0];
            default: return pc[0] * xs[1] + pc[1] * xs[2];
        }
    }
    // coefficient of xs[c] in row r (c: 0 n, 1 v, 2 delta) and of u_a,k
    __device__ __forceinline__ static double row_coef(int r, int c, const double* pc) {
        switch (r) {
            case 0: return c == 1 ? 1.0 : 0.0;
            case 1: return c == 2 ? 1.0 : 0.0;
            case 2: return c == 0 ? 1.0 : 0.0;
            default: return c == 1 ? pc[0] : (c == 2 ? pc[1] : 0.0);
        }
    }
    __device__ __forceinline__ static double row_ucoef(int r, int uc, const double* pc) { return 0.0; }

    // bounds of row r at step k given the free response xf (= A_bar x0 + d_bar at step k),
    // the linearisation point and g0  (kinematic_state_constraints.m:29-39,
    // kinematic_tyre_linearise_constraints.m:30-32)
    __device__ static void row_bounds(int r, const double* xf, const double* xl, const double* ul,
                                      const double* pc, const double* g0, const fsae_params& p,
                                      double& lo, double& up) {
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

}  // namespace fsae
