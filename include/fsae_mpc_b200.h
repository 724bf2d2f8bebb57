/* fsae_mpc_b200 -- C-ABI of the B200-native batched LTV-MPC step.
 *
 * Drop-in boundary for kerry-he/fsae-mpc's per-timestep LTV-MPC path.  Every entry
 * point names the reference interface it replaces (file:line in the reference tree).
 * Plain pointers and sizes only; no torch / C++ types.  All matrices use MATLAB's
 * column-major layout with the BATCH as the trailing dimension, so a MEX shim can
 * pass mxGetPr() pointers straight through.
 *
 *   x0     [N_x        x B]      x_ref [N_x x N_steps x B]
 *   x_lin  [N_x x N_steps x B]   u_lin [N_u x N_steps x B]
 *   u_opt  [N_u*N_steps x B]     x_opt [N_x*N_steps x B]
 *
 * "_host" entry points take host pointers (pageable or pinned) and do the H2D/D2H
 * copies themselves; "_dev" entry points take device pointers and a cudaStream_t
 * (passed as void*) and never synchronise.
 *
 * There is NO CPU fallback: every compute call returns FSAE_ERR_CUDA when no sm_100
 * device is usable.
 */
#ifndef FSAE_MPC_B200_H
#define FSAE_MPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSAE_OK 0
#define FSAE_ERR_ARG -1      /* bad argument (NULL pointer, unsupported horizon, ...) */
#define FSAE_ERR_CUDA -2     /* CUDA runtime error; see fsae_last_error() */
#define FSAE_ERR_UNSUPPORTED -3

#define FSAE_MAX_HORIZON 80   /* any N_steps in [1, 80], both models; N_steps = length(x_ref) as in
                                 ltvmpc_kinetmatic_curvilinear.m:17 -- not a compile-time choice of the caller */
#define FSAE_MAX_TRACKS 16
#define FSAE_MAX_PARAM_SETS 64

/* Model selectors (main.m:26 MODEL = "KINEMATIC" | "DYNAMIC"). */
#define FSAE_MODEL_KINEMATIC 0
#define FSAE_MODEL_DYNAMIC 1

/* Linearisation scheme: mpc/ltv/{kinematic,dynamic}/{euler,rk2,rk4}_*_curvilinear.m */
#define FSAE_LIN_EULER 1
#define FSAE_LIN_RK2 2
#define FSAE_LIN_RK4 4

/* exitflag values follow optimizers/matlab/qpOASES/qpOASES.m:44-48 */
#define FSAE_EXIT_SOLVED 0
#define FSAE_EXIT_MAXITER 1
#define FSAE_EXIT_INTERNAL -1
#define FSAE_EXIT_INFEASIBLE -2

/* Everything the reference hard-codes inside its .m files, gathered so that a
 * parameter sweep can vary it per problem.  fsae_default_params() fills in the
 * reference's constants (cited per field). */
typedef struct fsae_params {
    /* vehicle (f_curv_kin.m:13-14, f_curv_dyn.m:13-18, A_curv_dyn.m:43-46) */
    double lr, lf, mass, inertia, grav;
    double pac_B, pac_C, pac_D, pac_E;
    /* cost (ltvmpc_kinetmatic_curvilinear.m:32-35, ltvmpc_dynamic_curvilinear.m:32-35) */
    double Q[7];            /* first N_x entries used */
    double Q_terminal[7];
    double R[2];
    double R_soft[4];       /* first N_soft entries used (1 kinematic, 4 dynamic) */
    /* bounds (ltvmpc_*_curvilinear.m:23-29) */
    double u_lb[2], u_ub[2];
    double vel_lb, vel_ub;          /* state_idx(1): v / x_d   [0, inf]      */
    double delta_lb, delta_ub;      /* state_idx(2): delta     [-0.4, 0.4]   */
    double n_lb, n_ub;              /* soft_idx: n             [-0.75, 0.75] */
    double soft_far;                /* the 1e10 "no bound" of *_state_constraints.m:38-39 */
    /* kinematic tyre constraint (kinematic_tyre_linearise_constraints.m:31-32) */
    double ay_max;                  /* 5.0 */
    /* dynamic constraints (dynamic_slip_linearise_constraints.m:43-44,
       dynamic_tyre_linearise_constraints.m:5-6,18) */
    double slip_max;                /* 0.1 */
    double ac_max, al_max;          /* 9.163, 10.0 */
    /* solver */
    int lin_scheme;                 /* FSAE_LIN_*; reference: RK2 kinematic, RK4 dynamic */
    int max_iter;                   /* active-set iteration cap -> exitflag 1; <= 0: qpOASES's
                                       heuristic 5*(nV+nC) (qpOASES_options.m:40-41) */
    double feas_tol;                /* constraint violation tolerance (absolute) */
    double flat_eps;                /* curvature given to zero-Hessian slack variables when
                                       one of their bounds leaves the working set */
} fsae_params;

typedef struct fsae_ctx fsae_ctx;   /* opaque */

/* ---- context ------------------------------------------------------------------ */
int fsae_create(fsae_ctx** ctx, int device);
int fsae_destroy(fsae_ctx* ctx);
const char* fsae_last_error(const fsae_ctx* ctx);
const char* fsae_version(void);
/* number of kernels this library launched on ctx since creation (bench: gpu_launches) */
int64_t fsae_launch_count(const fsae_ctx* ctx);
/* device time (ms) of the most recent fused-step kernel, measured with CUDA events on the
 * context's own stream; 0 if none.  Only valid after a _host call or a stream sync. */
float fsae_last_kernel_ms(const fsae_ctx* ctx);
void* fsae_stream(const fsae_ctx* ctx);   /* the cudaStream_t the _host calls run on */

void fsae_default_params(int model, fsae_params* p);
/* Install parameter set `id` (0 is pre-installed with the reference defaults). */
int fsae_set_params(fsae_ctx* ctx, int id, const fsae_params* p);

/* Track = what main.m:15-18 builds: arclength-parameterised Bezier spline
 * coefficients x_spline/y_spline [n_seg x 4] (MATLAB column-major) and dl.
 * Replaces the `kappa` anonymous function argument of ltvmpc_*_curvilinear
 * (main.m:18, spline/interpolate_curvature.m:1). */
int fsae_set_track(fsae_ctx* ctx, int track_id, const double* x_spline, const double* y_spline,
                   int n_seg, double dl);

/* ---- util/obtain_reference.m:1-50 (call site main.m:115), batched over vehicles ----------
 * Re-parameterises a planned lap from arclength to time: plan_x [8 x N_s] column-major
 * (n, mu, x_d, y_d, theta_d, delta, a, delta_d at every ds metres -- the `x` the reference's
 * minimum-time planner returns), plan_t [N_s] the time spent in each segment.  Vehicle b starts at
 * arclength s0[b]; x_ref [7 x N_t x B] (s, n, mu, x_d, y_d, theta_d, delta at dt, 2dt, .. N_t dt) is
 * the `x_ref` argument of ltvmpc_dynamic_curvilinear.  Bit-identical to the .m file.  Host pointers. */
int fsae_obtain_reference_host(fsae_ctx* ctx, const double* plan_x, const double* plan_t, int N_s, double ds,
                               const double* s0, int B, double dt, int N_t, double* x_ref);

/* ---- spline/interpolate_curvature.m:1-20, batched over s (host pointers) -------- */
int fsae_interpolate_curvature_host(fsae_ctx* ctx, int track_id, const double* s, int64_t n,
                                    double* kappa_out);

/* ---- mpc/ltv/kinematic/rk2_kinematic_curvilinear.m:1 (and euler/rk4, dynamic/...) --
 * A [N_x x N_x x N_steps x B], Bm [N_x x N_u x N_steps x B], d [N_x x N_steps x B]:
 * the CONTINUOUS-time linearisation the reference functions return. */
int fsae_linearise_host(fsae_ctx* ctx, int model, int B, int N_steps, double dt,
                        const int32_t* track_id, const int32_t* param_id,
                        const double* x_lin, const double* u_lin,
                        double* A, double* Bm, double* d);

/* ---- sequential_integration.m + *_state_constraints.m + generate_qp.m ------------
 * Emits the condensed QP exactly as the reference hands it to qpOASES
 * (ltvmpc_kinetmatic_curvilinear.m:38-41 / ltvmpc_dynamic_curvilinear.m:38-41):
 *   H [nV x nV x B], f [nV x B], xA [nC x nV x B], lbA/ubA [nC x B], lb/ub [nV x B],
 *   A_bar [N_x*N x N_x x B], B_bar [N_x*N x nV x B], d_bar [N_x*N x B], cost_const [B].
 * nV = N_u*N + N_soft; nC = 6*N (kinematic) or 20*N (dynamic).  Any output may be NULL. */
int fsae_condense_host(fsae_ctx* ctx, int model, int B, int N_steps, double dt,
                       const int32_t* track_id, const int32_t* param_id,
                       const double* x0, const double* x_ref,
                       const double* x_lin, const double* u_lin,
                       double* H, double* f, double* xA, double* lbA, double* ubA,
                       double* lb, double* ub,
                       double* A_bar, double* B_bar, double* d_bar, double* cost_const);

/* ---- the fused step ----------------------------------------------------------------
 * [u_opt, x_opt, QP, exitflag, fval, slack_opt] = ltvmpc_kinetmatic_curvilinear(
 *        x0, x_ref, kappa, dt, x_lin, u_lin, QP)      (ltvmpc_kinetmatic_curvilinear.m:1)
 * and ltvmpc_dynamic_curvilinear (ltvmpc_dynamic_curvilinear.m:1), for B independent
 * problems.  track_id / param_id may be NULL (all problems use id 0).
 * Optional outputs (may be NULL): iters [B] active-set iterations, workingSetB
 * [nV x B] / workingSetC [nC x B] as int8 in qpOASES's -1/0/+1 encoding
 * (qpOASES.m:52-60). */
int fsae_ltvmpc_host(fsae_ctx* ctx, int model, int B, int N_steps, double dt,
                     const int32_t* track_id, const int32_t* param_id,
                     const double* x0, const double* x_ref,
                     const double* x_lin, const double* u_lin,
                     double* u_opt, double* x_opt, int32_t* exitflag, double* fval,
                     double* slack_opt, int32_t* iters,
                     int8_t* workingSetB, int8_t* workingSetC);

int fsae_ltvmpc_dev(fsae_ctx* ctx, int model, int B, int N_steps, double dt,
                    const int32_t* track_id, const int32_t* param_id,
                    const double* x0, const double* x_ref,
                    const double* x_lin, const double* u_lin,
                    double* u_opt, double* x_opt, int32_t* exitflag, double* fval,
                    double* slack_opt, int32_t* iters,
                    int8_t* workingSetB, int8_t* workingSetC, void* stream);

/* Host-path selection of the _host entry points.  Pinned caller buffers (cudaHostAlloc / cudaHostRegister)
 * are copied directly.  PAGEABLE caller buffers -- what a MEX gateway gets from mxGetPr() -- go through a
 * pinned staging ring inside the library, filled and drained by helper threads (FSAE_COPY_THREADS, default
 * min(8, cores/2)), so that the PCIe copies stay asynchronous and overlap the kernels.
 * mode 0 = automatic (default), 1 = always direct copies, 2 = always the staging ring; returns the previous mode. */
int fsae_set_host_staging(fsae_ctx* ctx, int mode);
/* 1 if the most recent fsae_ltvmpc_host call used the staging ring, 0 if it copied directly. */
int fsae_last_host_path(const fsae_ctx* ctx);

/* ---- device pool: ONE host thread (e.g. one MATLAB process) drives all GPUs of the box ---------------
 * north_star: "the batch ... sharded embarrassingly across the 8 GPUs of one B200 box, with no NCCL on the
 * solve path".  A pool owns one context per device; fsae_ltvmpc_host_pool splits the batch into contiguous
 * shards (fsae_shard_range: sizes differ by at most one) and runs each through fsae_ltvmpc_host on its
 * device, concurrently.  Same arguments and results as fsae_ltvmpc_host.
 * devices == NULL: the first n_devices visible devices (all of them if n_devices <= 0). */
typedef struct fsae_pool fsae_pool;
int fsae_pool_create(fsae_pool** pool, const int* devices, int n_devices);
int fsae_pool_destroy(fsae_pool* pool);
int fsae_pool_size(const fsae_pool* pool);
fsae_ctx* fsae_pool_ctx(fsae_pool* pool, int i);          /* the context of the i-th device (diagnostics) */
const char* fsae_pool_last_error(const fsae_pool* pool);
int fsae_pool_set_track(fsae_pool* pool, int track_id, const double* x_spline, const double* y_spline,
                        int n_seg, double dl);
int fsae_pool_set_params(fsae_pool* pool, int id, const fsae_params* p);
void fsae_shard_range(int64_t total, int rank, int world, int64_t* lo, int64_t* hi);
int fsae_ltvmpc_host_pool(fsae_pool* pool, int model, int B, int N_steps, double dt,
                          const int32_t* track_id, const int32_t* param_id,
                          const double* x0, const double* x_ref,
                          const double* x_lin, const double* u_lin,
                          double* u_opt, double* x_opt, int32_t* exitflag, double* fval,
                          double* slack_opt, int32_t* iters,
                          int8_t* workingSetB, int8_t* workingSetC);

/* ---- sequential QP: n_sqp repeated relinearise + condense + QP passes per problem, each
 * pass linearising at the previous pass's (x_opt, u_opt) -- BASELINE.json configs[3]
 * ("mpc/nonlinear SQP: repeated relinearise+QP iterations per step").  It is exactly what
 * main.m:118-127 does across consecutive time steps (x_lin/u_lin = previous x_opt/u_opt),
 * iterated on a frozen x0/x_ref.  n_sqp = 1 is fsae_ltvmpc_host.  Outputs are those of the
 * last pass; exitflag is the first non-zero flag met (0 if every pass solved). */
int fsae_ltvmpc_sqp_host(fsae_ctx* ctx, int model, int B, int N_steps, double dt, int n_sqp,
                         const int32_t* track_id, const int32_t* param_id,
                         const double* x0, const double* x_ref,
                         const double* x_lin, const double* u_lin,
                         double* u_opt, double* x_opt, int32_t* exitflag, double* fval,
                         double* slack_opt, int32_t* iters);

/* ---- batch driver: main.m's closed loop (main.m:90-190) for B vehicles ----------------
 * Per simulation step: projection of the plant state onto the track
 * (cartesian_to_curvilinear.m / closest_point.m), x0 and speed-ramp reference (main.m:92-114),
 * the fused LTV-MPC step linearised at the previous prediction (main.m:118-127), then the
 * actuator PIDs and the Cartesian dynamic plant (main.m:166-179, pid_controller.m,
 * integrate_cart_dyn.m).  A vehicle stops when s >= track length (main.m:102).
 *   plant0 [7 x B]; x_opt0 [N_x*N x B], u_opt0 [N_u*N x B]: the initial guess of main.m:44-55.
 * Outputs: plant_final [7 x B], steps [B] MPC steps taken; optional histories
 * n_hist [n_sim x B] (lateral deviation, main.m:101), plant_hist [7 x n_sim x B] (x_history),
 * exit_hist [n_sim x B]. */
int fsae_closed_loop_host(fsae_ctx* ctx, int model, int B, int N_steps, double dt, int n_sim,
                          double target_vel, const int32_t* track_id, const int32_t* param_id,
                          const double* plant0, const double* x_opt0, const double* u_opt0,
                          double* plant_final, int32_t* steps,
                          double* n_hist, double* plant_hist, int32_t* exit_hist);

/* ---- [x,fval,exitflag,iter,lambda,auxOutput] = qpOASES(H,g,A,lb,ub,lbA,ubA) ----------
 * (optimizers/matlab/qpOASES/qpOASES.m:22) for B independent dense QPs of one shape.
 * H [nV x nV x B] symmetric, A [nC x nV x B] column-major.  lambda [ (nV+nC) x B ] in
 * qpOASES's sign convention.  Optional outputs may be NULL.  nV <= 191 (the condensed MPC QPs up to
 * horizon 80 have nV = 161 / 164); larger problems return FSAE_ERR_UNSUPPORTED. */
int fsae_qpoases_host(fsae_ctx* ctx, int B, int nV, int nC,
                      const double* H, const double* g, const double* A,
                      const double* lb, const double* ub, const double* lbA, const double* ubA,
                      double* x, double* fval, int32_t* exitflag, int32_t* iters,
                      double* lambda, int8_t* workingSetB, int8_t* workingSetC);

/* ---- measurement helpers (bench.py) -------------------------------------------------
 * Sum over all problems solved on ctx so far of the active-set events
 * out3 = {constraints added, constraints dropped, refresh steps}; reset != 0 clears them. */
int fsae_debug_counters(fsae_ctx* ctx, uint64_t* out3, int reset);
/* Select the fused kernel variant: 2 (default) = register-tiled product kernel,
 * 1 = shared-memory variant kept as an in-library cross-check; 21 / 26 / 28 / 29 = other warp
 * counts and block sizes of the register-tiled kernel; 31 / 32 = the kinematic solver in integrator
 * coordinates with the column-lane / row-lane operator layout (tests, horizons 20 and 40; only in the
 * cross-check build of the library).  Returns the previous value. */
int fsae_debug_set_kernel_version(fsae_ctx* ctx, int version);
/* Debug taps of the fused kernel (tests): DEVICE buffers that the following fsae_ltvmpc_dev calls
 * fill, per problem and column-major, with the condensed Hessian H [nV x nV x B]
 * (generate_qp.m:29), the gradient g [nV x B] (generate_qp.m:30) and the initial operator of the
 * dual active-set solve M = [e_slack | J], J'HJ = I [nV x nV x B].  NULL switches a tap off. */
int fsae_debug_set_taps(fsae_ctx* ctx, double* d_H, double* d_g, double* d_M);
/* Measured FP64 FMA peak of the device in TFLOP/s (FMA = 2 flops): the roofline
 * denominator for the solve kernel, which MEASURED_PEAKS.json does not carry. */
int fsae_probe_fp64_tflops(fsae_ctx* ctx, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* FSAE_MPC_B200_H */
