"""bench.py's CPU-baseline / --impl reference helper: times the oracle's C restatement of the
reference's kinematic LTV-MPC step (oracle/ltvmpc_oracle.c, dense formulation like the MATLAB
code) on the host cores.  This is one of the few places allowed to execute oracle/ (as the
thing TIMED AS A BASELINE, never as part of the product path).

kind = "port": the reference itself is MATLAB + Windows-only qpOASES MEX binaries and cannot
run on this box.  A compiled C port is considerably faster than interpreted MATLAB, so the
baseline is conservative (flatters the CPU)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(ROOT, "oracle", "libltvmpc_oracle.so")


def load():
    if not os.path.exists(LIB):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    lib = C.CDLL(LIB)
    dp, ip, bp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int8)
    lib.oracle_ltvmpc_kinematic_batch.restype = C.c_int
    lib.oracle_ltvmpc_kinematic_batch.argtypes = [dp, dp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_double,
                                                  dp, dp, dp, dp, dp, dp, ip, dp, dp, ip, C.c_int]
    lib.oracle_ltvmpc_dynamic_batch.restype = C.c_int
    lib.oracle_ltvmpc_dynamic_batch.argtypes = lib.oracle_ltvmpc_kinematic_batch.argtypes
    lib.oracle_ltvmpc_kinematic.restype = C.c_int
    lib.oracle_ltvmpc_kinematic.argtypes = [dp, dp, C.c_int, C.c_double, C.c_int, C.c_double,
                                            dp, dp, dp, dp, dp, dp, ip, dp, dp, ip, bp, bp]
    return lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Baseline:
    kind = "port"

    def __init__(self, track, threads=0):
        self.lib = load()
        self.xs = np.asfortranarray(track[0], dtype=np.float64)
        self.ys = np.asfortranarray(track[1], dtype=np.float64)
        self.dl = float(track[2])
        # all host cores the process may use, explicitly: torchrun exports OMP_NUM_THREADS=1
        if not threads:
            try:
                threads = len(os.sched_getaffinity(0))
            except AttributeError:
                threads = os.cpu_count() or 1
        self.threads = threads
        self.cores = threads
        self.last = None

    NX, NS, FN = 5, 1, "oracle_ltvmpc_kinematic_batch"

    def run(self, x0, x_ref, x_lin, u_lin, dt):
        """Solve all problems of the (C-ABI layout) batch; returns the number solved."""
        B, N = x_ref.shape[0], x_ref.shape[1]
        assert x0.shape[0] == B == x_lin.shape[0] == u_lin.shape[0], "all four arrays must hold the same problems"
        x0, x_ref, x_lin, u_lin = (np.ascontiguousarray(a, dtype=np.float64) for a in (x0, x_ref, x_lin, u_lin))
        out = dict(u_opt=np.empty((B, 2 * N)), x_opt=np.empty((B, self.NX * N)), exitflag=np.empty(B, np.int32),
                   fval=np.empty(B), slack=np.empty((B, self.NS)), iters=np.empty(B, np.int32))
        used = getattr(self.lib, self.FN)(
            _dp(self.xs), _dp(self.ys), self.xs.shape[0], self.dl, B, N, float(dt),
            _dp(x0), _dp(x_ref), _dp(x_lin), _dp(u_lin), _dp(out["u_opt"]), _dp(out["x_opt"]),
            out["exitflag"].ctypes.data_as(C.POINTER(C.c_int32)), _dp(out["fval"]), _dp(out["slack"]),
            out["iters"].ctypes.data_as(C.POINTER(C.c_int32)), int(self.threads))
        self.cores = used
        self.last = out
        return B


class DynamicBaseline(Baseline):
    """The dynamic (tyre-force) model, ltvmpc_dynamic_curvilinear.m:1 -- oracle_ltvmpc_dynamic_batch."""
    NX, NS, FN = 7, 4, "oracle_ltvmpc_dynamic_batch"
