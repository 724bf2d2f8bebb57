"""Host-side mirror of the reference's LTV-MPC call surface, batched, on one B200.

The reference host language is MATLAB (not available in this image); this module is the
Python twin of the MEX shim described in INTEGRATION.md and keeps the reference's
function names, argument meaning and outputs:

    reference (MATLAB, one problem)                       here (B problems)
    ---------------------------------------------------------------------------------
    kappa = @(s) interpolate_curvature(s, xs, ys, dl)     mpc.set_track(id, xs, ys, dl)
    rk2_kinematic_curvilinear(x, u, kappa, dt)            mpc.linearise(KINEMATIC, x, u, dt)
    sequential_integration + *_state_constraints
        + generate_qp                                     mpc.condense(KINEMATIC, ...)
    ltvmpc_kinetmatic_curvilinear(x0,x_ref,kappa,dt,
                                  x_lin,u_lin,QP)         mpc.ltvmpc_kinetmatic_curvilinear(...)
    ltvmpc_dynamic_curvilinear(...)                       mpc.ltvmpc_dynamic_curvilinear(...)

Array layout: every per-problem matrix is passed in the C-ABI's memory layout, which is
MATLAB's column-major with the batch trailing.  As C-contiguous numpy arrays that reads
    x0 (B, N_x)   x_ref, x_lin (B, N_steps, N_x)   u_lin (B, N_steps, N_u)
    u_opt (B, N_steps*N_u)   x_opt (B, N_steps*N_x)    [== the reference's stacked vectors]

No CPU fallback: constructing FsaeMpc without the built CUDA library or without an
sm_100 GPU raises.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import Params, MODEL_KINEMATIC, MODEL_DYNAMIC

KINEMATIC = MODEL_KINEMATIC
DYNAMIC = MODEL_DYNAMIC
_DIMS = {KINEMATIC: (5, 2, 1), DYNAMIC: (7, 2, 4)}


class FsaeError(RuntimeError):
    pass


@dataclass
class MpcResult:
    """Outputs of ltvmpc_*_curvilinear.m:1 for B problems (+ solver diagnostics)."""
    u_opt: np.ndarray        # (B, N_u*N)
    x_opt: np.ndarray        # (B, N_x*N)
    exitflag: np.ndarray     # (B,) int32, qpOASES convention
    fval: np.ndarray         # (B,)
    slack_opt: np.ndarray    # (B, N_soft)
    iters: np.ndarray        # (B,) int32 active-set iterations
    workingSetB: np.ndarray  # (B, nV) int8
    workingSetC: np.ndarray  # (B, nC) int8


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))


def _f64(a, shape):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.shape != shape:
        raise ValueError(f"expected shape {shape}, got {a.shape}")
    return a


def default_params(model=KINEMATIC):
    p = Params()
    _lib.load().fsae_default_params(model, C.byref(p))
    return p


class FsaeMpc:
    """One context = one GPU + one stream + its track / parameter tables."""

    def __init__(self, device=0, lib_path=None):
        self._lib = _lib.load(lib_path)
        self._ctx = C.c_void_p()
        rc = self._lib.fsae_create(C.byref(self._ctx), int(device))
        if rc != 0:
            self._ctx = None
            raise FsaeError(
                f"fsae_create(device={device}) failed with {rc}: an sm_100 (B200) GPU is required; "
                "there is no CPU fallback")
        self.device = int(device)

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.fsae_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            msg = self._lib.fsae_last_error(self._ctx)
            raise FsaeError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    # ------------------------------------------------------------------ tables
    def set_track(self, track_id, x_spline, y_spline, dl):
        """main.m:15-18: the arclength spline behind `kappa`.  x_spline/y_spline (n_seg, 4)."""
        xs = np.asfortranarray(x_spline, dtype=np.float64)
        ys = np.asfortranarray(y_spline, dtype=np.float64)
        if xs.ndim != 2 or xs.shape[1] != 4 or ys.shape != xs.shape:
            raise ValueError("spline coefficient matrices must be (n_seg, 4)")
        self._check(self._lib.fsae_set_track(self._ctx, int(track_id), _dp(xs), _dp(ys),
                                             xs.shape[0], float(dl)), "fsae_set_track")

    def set_params(self, param_id, params):
        self._check(self._lib.fsae_set_params(self._ctx, int(param_id), C.byref(params)), "fsae_set_params")

    # ------------------------------------------------------------------ stages
    def interpolate_curvature(self, s, track_id=0):
        """spline/interpolate_curvature.m:1."""
        s = np.ascontiguousarray(s, dtype=np.float64).reshape(-1)
        out = np.empty_like(s)
        self._check(self._lib.fsae_interpolate_curvature_host(self._ctx, int(track_id), _dp(s), s.size, _dp(out)),
                    "fsae_interpolate_curvature_host")
        return out

    def _ids(self, B, track_id, param_id):
        t = None if track_id is None else np.ascontiguousarray(track_id, dtype=np.int32).reshape(B)
        p = None if param_id is None else np.ascontiguousarray(param_id, dtype=np.int32).reshape(B)
        return t, p

    def linearise(self, model, x_lin, u_lin, dt, track_id=None, param_id=None):
        """{euler,rk2,rk4}_{kinematic,dynamic}_curvilinear.m (scheme = params.lin_scheme).
        Returns A (B,N,NX,NX), Bm (B,N,NX,NU), d (B,N,NX) with A[b,k] the k-th matrix."""
        NX, NU, _ = _DIMS[model]
        x_lin = np.ascontiguousarray(x_lin, dtype=np.float64)
        B, N = x_lin.shape[0], x_lin.shape[1]
        x_lin = _f64(x_lin, (B, N, NX))
        u_lin = _f64(u_lin, (B, N, NU))
        t, p = self._ids(B, track_id, param_id)
        A = np.empty((B, N, NX, NX))
        Bm = np.empty((B, N, NU, NX))
        d = np.empty((B, N, NX))
        self._check(self._lib.fsae_linearise_host(self._ctx, model, B, N, float(dt), _ip(t), _ip(p),
                                                  _dp(x_lin), _dp(u_lin), _dp(A), _dp(Bm), _dp(d)),
                    "fsae_linearise_host")
        return A.transpose(0, 1, 3, 2), Bm.transpose(0, 1, 3, 2), d

    def obtain_reference(self, plan_x, ds, plan_t, s0, dt, N_t):
        """util/obtain_reference.m for a batch of vehicles: plan_x (N_s, 8) [the planner's x, one row per
        arclength sample], plan_t (N_s,), s0 (B,) -> x_ref (B, N_t, 7) in the C-ABI layout."""
        plan_x = np.ascontiguousarray(plan_x, dtype=np.float64).reshape(-1)
        plan_t = np.ascontiguousarray(plan_t, dtype=np.float64).reshape(-1)
        s0 = np.ascontiguousarray(np.atleast_1d(s0), dtype=np.float64)
        N_s, B = plan_t.size, s0.size
        if plan_x.size != 8 * N_s:
            raise ValueError("plan_x must hold 8 values per arclength sample")
        out = np.empty((B, int(N_t), 7), dtype=np.float64)
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        self._check(self._lib.fsae_obtain_reference_host(self._ctx, dp(plan_x), dp(plan_t), N_s, float(ds), dp(s0), B,
                                                         float(dt), int(N_t), dp(out)), "fsae_obtain_reference_host")
        return out

    def condense(self, model, x0, x_ref, dt, x_lin, u_lin, track_id=None, param_id=None):
        """sequential_integration.m + *_state_constraints.m + generate_qp.m.  Returns a dict of
        per-problem matrices in natural (row, col) indexing."""
        NX, NU, NS = _DIMS[model]
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        B = x0.shape[0]
        N = np.asarray(x_ref).shape[1]
        x0 = _f64(x0, (B, NX))
        x_ref = _f64(x_ref, (B, N, NX))
        x_lin = _f64(x_lin, (B, N, NX))
        u_lin = _f64(u_lin, (B, N, NU))
        t, p = self._ids(B, track_id, param_id)
        nV, nC, nXN = NU * N + NS, (6 if model == KINEMATIC else 20) * N, NX * N
        o = dict(H=np.empty((B, nV, nV)), f=np.empty((B, nV)), xA=np.empty((B, nV, nC)),
                 lbA=np.empty((B, nC)), ubA=np.empty((B, nC)), lb=np.empty((B, nV)), ub=np.empty((B, nV)),
                 A_bar=np.empty((B, NX, nXN)), B_bar=np.empty((B, nV, nXN)), d_bar=np.empty((B, nXN)),
                 const=np.empty(B))
        self._check(self._lib.fsae_condense_host(
            self._ctx, model, B, N, float(dt), _ip(t), _ip(p), _dp(x0), _dp(x_ref), _dp(x_lin), _dp(u_lin),
            _dp(o["H"]), _dp(o["f"]), _dp(o["xA"]), _dp(o["lbA"]), _dp(o["ubA"]), _dp(o["lb"]), _dp(o["ub"]),
            _dp(o["A_bar"]), _dp(o["B_bar"]), _dp(o["d_bar"]), _dp(o["const"])), "fsae_condense_host")
        for k in ("H", "xA", "A_bar", "B_bar"):
            o[k] = o[k].transpose(0, 2, 1)       # column-major blocks -> (row, col)
        return o

    # ------------------------------------------------------------------ the fused step
    def _ltvmpc(self, model, x0, x_ref, dt, x_lin, u_lin, track_id, param_id):
        NX, NU, NS = _DIMS[model]
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        B = x0.shape[0]
        N = np.asarray(x_ref).shape[1]
        x0 = _f64(x0, (B, NX))
        x_ref = _f64(x_ref, (B, N, NX))
        x_lin = _f64(x_lin, (B, N, NX))
        u_lin = _f64(u_lin, (B, N, NU))
        t, p = self._ids(B, track_id, param_id)
        nV, nC = NU * N + NS, (6 if model == KINEMATIC else 20) * N
        r = MpcResult(np.empty((B, NU * N)), np.empty((B, NX * N)), np.empty(B, np.int32), np.empty(B),
                      np.empty((B, NS)), np.empty(B, np.int32), np.empty((B, nV), np.int8),
                      np.empty((B, nC), np.int8))
        self._check(self._lib.fsae_ltvmpc_host(
            self._ctx, model, B, N, float(dt), _ip(t), _ip(p), _dp(x0), _dp(x_ref), _dp(x_lin), _dp(u_lin),
            _dp(r.u_opt), _dp(r.x_opt), _ip(r.exitflag), _dp(r.fval), _dp(r.slack_opt), _ip(r.iters),
            r.workingSetB.ctypes.data_as(C.POINTER(C.c_int8)),
            r.workingSetC.ctypes.data_as(C.POINTER(C.c_int8))), "fsae_ltvmpc_host")
        return r

    def ltvmpc_kinetmatic_curvilinear(self, x0, x_ref, dt, x_lin, u_lin, track_id=None, param_id=None):
        """mpc/ltv/kinematic/ltvmpc_kinetmatic_curvilinear.m:1 (name kept, typo included)."""
        return self._ltvmpc(KINEMATIC, x0, x_ref, dt, x_lin, u_lin, track_id, param_id)

    def ltvmpc_dynamic_curvilinear(self, x0, x_ref, dt, x_lin, u_lin, track_id=None, param_id=None):
        """mpc/ltv/dynamic/ltvmpc_dynamic_curvilinear.m:1."""
        return self._ltvmpc(DYNAMIC, x0, x_ref, dt, x_lin, u_lin, track_id, param_id)

    def qpOASES(self, H, g, A, lb, ub, lbA, ubA):
        """[x,fval,exitflag,iter,lambda,auxOutput] = qpOASES(H,g,A,lb,ub,lbA,ubA)
        (optimizers/matlab/qpOASES/qpOASES.m:22) for B dense QPs of one shape.
        H (B,nV,nV) symmetric, g (B,nV), A (B,nC,nV), bounds (B,nV)/(B,nC).  Returns a dict."""
        H = np.ascontiguousarray(H, dtype=np.float64)
        B, nV = H.shape[0], H.shape[1]
        A = np.zeros((B, 0, nV)) if A is None else np.asarray(A, dtype=np.float64)
        nC = A.shape[1]
        At = np.ascontiguousarray(A.transpose(0, 2, 1))          # column-major [nC x nV] per problem
        g, lb, ub = (_f64(v, (B, nV)) for v in (g, lb, ub))
        lbA = _f64(lbA if nC else np.zeros((B, 0)), (B, nC))
        ubA = _f64(ubA if nC else np.zeros((B, 0)), (B, nC))
        o = dict(x=np.empty((B, nV)), fval=np.empty(B), exitflag=np.empty(B, np.int32), iter=np.empty(B, np.int32),
                 lam=np.empty((B, nV + nC)), workingSetB=np.empty((B, nV), np.int8),
                 workingSetC=np.empty((B, max(nC, 1)), np.int8))
        bp = lambda a: a.ctypes.data_as(C.POINTER(C.c_int8))
        self._check(self._lib.fsae_qpoases_host(
            self._ctx, B, nV, nC, _dp(H), _dp(g), _dp(At), _dp(lb), _dp(ub), _dp(lbA), _dp(ubA),
            _dp(o["x"]), _dp(o["fval"]), _ip(o["exitflag"]), _ip(o["iter"]), _dp(o["lam"]),
            bp(o["workingSetB"]), bp(o["workingSetC"])), "fsae_qpoases_host")
        o["workingSetC"] = o["workingSetC"][:, :nC]
        return o

    def ltvmpc_sqp(self, model, x0, x_ref, dt, x_lin, u_lin, n_sqp, track_id=None, param_id=None):
        """n_sqp repeated relinearise + QP passes on a frozen (x0, x_ref): fsae_ltvmpc_sqp_host."""
        NX, NU, NS = _DIMS[model]
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        B = x0.shape[0]
        N = np.asarray(x_ref).shape[1]
        x0 = _f64(x0, (B, NX))
        x_ref = _f64(x_ref, (B, N, NX))
        x_lin = _f64(x_lin, (B, N, NX))
        u_lin = _f64(u_lin, (B, N, NU))
        t, p = self._ids(B, track_id, param_id)
        r = MpcResult(np.empty((B, NU * N)), np.empty((B, NX * N)), np.empty(B, np.int32), np.empty(B),
                      np.empty((B, NS)), np.empty(B, np.int32), np.zeros((B, 0), np.int8), np.zeros((B, 0), np.int8))
        self._check(self._lib.fsae_ltvmpc_sqp_host(
            self._ctx, model, B, N, float(dt), int(n_sqp), _ip(t), _ip(p), _dp(x0), _dp(x_ref), _dp(x_lin), _dp(u_lin),
            _dp(r.u_opt), _dp(r.x_opt), _ip(r.exitflag), _dp(r.fval), _dp(r.slack_opt), _ip(r.iters)),
            "fsae_ltvmpc_sqp_host")
        return r

    def closed_loop(self, model, plant0, n_sim, N=40, dt=0.05, target_vel=20.0, x_opt0=None, u_opt0=None,
                    track_id=None, param_id=None, history=True):
        """main.m's closed loop for B vehicles (fsae_closed_loop_host).  plant0 (B,7).  The default
        initial guess is main.m:44-55 (quadratic arc length, linear speed, constant 10 m/s^2).
        Returns dict(plant, steps, n_hist (B,n_sim), plant_hist (B,n_sim,7), exit_hist)."""
        NX, NU, NS = _DIMS[model]
        plant0 = np.ascontiguousarray(plant0, dtype=np.float64)
        B = plant0.shape[0]
        plant0 = _f64(plant0, (B, 7))
        if x_opt0 is None:
            t = dt * np.arange(1, N + 1)
            x1 = np.zeros((N, NX)); x1[:, 0] = 10 * t ** 2 / 2; x1[:, 3] = 10 * t
            u1 = np.zeros((N, NU)); u1[:, 0] = 10
            x_opt0 = np.tile(x1[None], (B, 1, 1)); u_opt0 = np.tile(u1[None], (B, 1, 1))
        x_opt0 = _f64(x_opt0, (B, N, NX)); u_opt0 = _f64(u_opt0, (B, N, NU))
        t_, p_ = self._ids(B, track_id, param_id)
        o = dict(plant=np.empty((B, 7)), steps=np.empty(B, np.int32))
        if history:
            o.update(n_hist=np.empty((B, n_sim)), plant_hist=np.empty((B, n_sim, 7)), exit_hist=np.empty((B, n_sim), np.int32))
        self._check(self._lib.fsae_closed_loop_host(
            self._ctx, model, B, int(N), float(dt), int(n_sim), float(target_vel), _ip(t_), _ip(p_),
            _dp(plant0), _dp(x_opt0), _dp(u_opt0), _dp(o["plant"]), _ip(o["steps"]),
            _dp(o["n_hist"]) if history else None, _dp(o["plant_hist"]) if history else None,
            _ip(o["exit_hist"]) if history else None), "fsae_closed_loop_host")
        return o

    def ltvmpc_dev(self, model, B, N, dt, ptrs, stream=0):
        """Device-pointer entry (fsae_ltvmpc_dev).  `ptrs` is a dict of integer device
        addresses: x0,x_ref,x_lin,u_lin,u_opt,x_opt,exitflag,fval,slack_opt and optionally
        track_id,param_id,iters,workingSetB,workingSetC.  Never synchronises."""
        g = lambda k: C.c_void_p(ptrs.get(k) or None)
        self._check(self._lib.fsae_ltvmpc_dev(
            self._ctx, model, int(B), int(N), float(dt), g("track_id"), g("param_id"),
            g("x0"), g("x_ref"), g("x_lin"), g("u_lin"), g("u_opt"), g("x_opt"), g("exitflag"),
            g("fval"), g("slack_opt"), g("iters"), g("workingSetB"), g("workingSetC"),
            C.c_void_p(stream or None)), "fsae_ltvmpc_dev")

    # ------------------------------------------------------------------ diagnostics
    @property
    def launch_count(self):
        return int(self._lib.fsae_launch_count(self._ctx))

    @property
    def last_kernel_ms(self):
        return float(self._lib.fsae_last_kernel_ms(self._ctx))

    @property
    def stream(self):
        return int(self._lib.fsae_stream(self._ctx) or 0)

    def set_taps(self, d_H=0, d_g=0, d_M=0):
        """Device addresses (0 = off) that the next ltvmpc_dev calls fill with H, g and the initial
        operator M = [e_slack | J] of every problem (fsae_debug_set_taps; tests)."""
        self._check(self._lib.fsae_debug_set_taps(self._ctx, C.c_void_p(d_H or None), C.c_void_p(d_g or None),
                                                  C.c_void_p(d_M or None)), "fsae_debug_set_taps")

    def set_host_staging(self, mode):
        """Host path of the _host calls: 0 automatic (pinned staging ring when a caller buffer is pageable),
        1 always direct copies, 2 always the ring.  Returns the previous mode."""
        return int(self._lib.fsae_set_host_staging(self._ctx, int(mode)))

    @property
    def last_host_path(self):
        """1 if the last fused-step host call went through the pinned staging ring, else 0."""
        return int(self._lib.fsae_last_host_path(self._ctx))

    def set_kernel_version(self, v):
        """2 = register-tiled product kernel (default), 1 = shared-memory cross-check variant."""
        return int(self._lib.fsae_debug_set_kernel_version(self._ctx, int(v)))

    def probe_fp64_tflops(self):
        out = C.c_double()
        self._check(self._lib.fsae_probe_fp64_tflops(self._ctx, C.byref(out)), "fsae_probe_fp64_tflops")
        return float(out.value)

    def counters(self, reset=False):
        """(adds, drops, refreshes) summed over all problems solved so far."""
        out = (C.c_uint64 * 3)()
        self._check(self._lib.fsae_debug_counters(self._ctx, out, int(reset)), "fsae_debug_counters")
        return tuple(int(v) for v in out)


class FsaePool:
    """One host thread drives every GPU of the box (fsae_pool_*): the batch is split into contiguous shards
    (fsae_shard_range == sharding.shard_range) and each shard runs through fsae_ltvmpc_host on its own device,
    concurrently.  No collective: the problems are independent.  This is the path a single MATLAB process uses
    (matlab/fsae_mpc_b200_handle.m); bench.py's multi-GPU arm uses one process per GPU instead (torchrun)."""

    def __init__(self, devices=None, lib_path=None):
        self._lib = _lib.load(lib_path)
        self._pool = C.c_void_p()
        if devices is None:
            rc = self._lib.fsae_pool_create(C.byref(self._pool), None, 0)
        else:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            rc = self._lib.fsae_pool_create(C.byref(self._pool), arr, len(devices))
        if rc != 0:
            self._pool = None
            raise FsaeError(f"fsae_pool_create failed with {rc}: sm_100 (B200) GPUs are required; there is no CPU fallback")

    def close(self):
        if getattr(self, "_pool", None):
            self._lib.fsae_pool_destroy(self._pool)
            self._pool = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return int(self._lib.fsae_pool_size(self._pool))

    def _check(self, rc, what):
        if rc != 0:
            msg = self._lib.fsae_pool_last_error(self._pool)
            raise FsaeError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def set_track(self, track_id, x_spline, y_spline, dl):
        xs = np.asfortranarray(x_spline, dtype=np.float64)
        ys = np.asfortranarray(y_spline, dtype=np.float64)
        self._check(self._lib.fsae_pool_set_track(self._pool, int(track_id), _dp(xs), _dp(ys), xs.shape[0], float(dl)),
                    "fsae_pool_set_track")

    def set_params(self, param_id, params):
        self._check(self._lib.fsae_pool_set_params(self._pool, int(param_id), C.byref(params)), "fsae_pool_set_params")

    def shard_range(self, total, rank):
        lo, hi = C.c_int64(), C.c_int64()
        self._lib.fsae_shard_range(int(total), int(rank), len(self), C.byref(lo), C.byref(hi))
        return int(lo.value), int(hi.value)

    def ltvmpc(self, model, x0, x_ref, dt, x_lin, u_lin, track_id=None, param_id=None):
        """ltvmpc_*_curvilinear for B problems over all devices of the pool (fsae_ltvmpc_host_pool)."""
        NX, NU, NS = _DIMS[model]
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        B = x0.shape[0]
        N = np.asarray(x_ref).shape[1]
        x0 = _f64(x0, (B, NX)); x_ref = _f64(x_ref, (B, N, NX)); x_lin = _f64(x_lin, (B, N, NX)); u_lin = _f64(u_lin, (B, N, NU))
        t = None if track_id is None else np.ascontiguousarray(track_id, dtype=np.int32)
        p = None if param_id is None else np.ascontiguousarray(param_id, dtype=np.int32)
        nV, nC = NU * N + NS, (6 if model == KINEMATIC else 20) * N
        r = MpcResult(np.empty((B, NU * N)), np.empty((B, NX * N)), np.empty(B, np.int32), np.empty(B),
                      np.empty((B, NS)), np.empty(B, np.int32), np.empty((B, nV), np.int8), np.empty((B, nC), np.int8))
        self._check(self._lib.fsae_ltvmpc_host_pool(
            self._pool, model, B, N, float(dt), _ip(t), _ip(p), _dp(x0), _dp(x_ref), _dp(x_lin), _dp(u_lin),
            _dp(r.u_opt), _dp(r.x_opt), _ip(r.exitflag), _dp(r.fval), _dp(r.slack_opt), _ip(r.iters),
            r.workingSetB.ctypes.data_as(C.POINTER(C.c_int8)), r.workingSetC.ctypes.data_as(C.POINTER(C.c_int8))),
            "fsae_ltvmpc_host_pool")
        return r
