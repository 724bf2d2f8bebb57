"""ctypes binding of the C-ABI in include/fsae_mpc_b200.h.  Loads the in-tree .so and
fails loudly if it is missing -- there is no CPU fallback."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfsae_mpc_b200.so")

FSAE_OK, FSAE_ERR_ARG, FSAE_ERR_CUDA, FSAE_ERR_UNSUPPORTED = 0, -1, -2, -3
MODEL_KINEMATIC, MODEL_DYNAMIC = 0, 1
LIN_EULER, LIN_RK2, LIN_RK4 = 1, 2, 4


class Params(C.Structure):
    """struct fsae_params (include/fsae_mpc_b200.h)."""
    _fields_ = [
        ("lr", C.c_double), ("lf", C.c_double), ("mass", C.c_double), ("inertia", C.c_double),
        ("grav", C.c_double),
        ("pac_B", C.c_double), ("pac_C", C.c_double), ("pac_D", C.c_double), ("pac_E", C.c_double),
        ("Q", C.c_double * 7), ("Q_terminal", C.c_double * 7), ("R", C.c_double * 2),
        ("R_soft", C.c_double * 4),
        ("u_lb", C.c_double * 2), ("u_ub", C.c_double * 2),
        ("vel_lb", C.c_double), ("vel_ub", C.c_double),
        ("delta_lb", C.c_double), ("delta_ub", C.c_double),
        ("n_lb", C.c_double), ("n_ub", C.c_double),
        ("soft_far", C.c_double),
        ("ay_max", C.c_double),
        ("slip_max", C.c_double),
        ("ac_max", C.c_double), ("al_max", C.c_double),
        ("lin_scheme", C.c_int), ("max_iter", C.c_int),
        ("feas_tol", C.c_double), ("flat_eps", C.c_double),
    ]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_int8)
_ctx = C.c_void_p

# every symbol include/fsae_mpc_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "fsae_create": (C.c_int, [C.POINTER(_ctx), C.c_int]),
    "fsae_destroy": (C.c_int, [_ctx]),
    "fsae_last_error": (C.c_char_p, [_ctx]),
    "fsae_version": (C.c_char_p, []),
    "fsae_launch_count": (C.c_int64, [_ctx]),
    "fsae_last_kernel_ms": (C.c_float, [_ctx]),
    "fsae_stream": (C.c_void_p, [_ctx]),
    "fsae_default_params": (None, [C.c_int, C.POINTER(Params)]),
    "fsae_set_params": (C.c_int, [_ctx, C.c_int, C.POINTER(Params)]),
    "fsae_set_track": (C.c_int, [_ctx, C.c_int, _dp, _dp, C.c_int, C.c_double]),
    "fsae_interpolate_curvature_host": (C.c_int, [_ctx, C.c_int, _dp, C.c_int64, _dp]),
    "fsae_linearise_host": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, C.c_double, _ip, _ip,
                                      _dp, _dp, _dp, _dp, _dp]),
    "fsae_condense_host": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, C.c_double, _ip, _ip,
                                     _dp, _dp, _dp, _dp] + [_dp] * 11),
    "fsae_ltvmpc_host": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, C.c_double, _ip, _ip,
                                   _dp, _dp, _dp, _dp, _dp, _dp, _ip, _dp, _dp, _ip, _bp, _bp]),
    "fsae_ltvmpc_sqp_host": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, _ip, _ip,
                                       _dp, _dp, _dp, _dp, _dp, _dp, _ip, _dp, _dp, _ip]),
    "fsae_closed_loop_host": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double, _ip, _ip,
                                        _dp, _dp, _dp, _dp, _ip, _dp, _dp, _ip]),
    "fsae_ltvmpc_dev": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
    "fsae_qpoases_host": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int] + [_dp] * 7
                          + [_dp, _dp, _ip, _ip, _dp, _bp, _bp]),
    "fsae_debug_counters": (C.c_int, [_ctx, C.POINTER(C.c_uint64), C.c_int]),
    "fsae_debug_set_kernel_version": (C.c_int, [_ctx, C.c_int]),
    "fsae_obtain_reference_host": (C.c_int, [_ctx, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_double,
                                             C.POINTER(C.c_double), C.c_int, C.c_double, C.c_int, C.POINTER(C.c_double)]),
    "fsae_debug_set_taps": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fsae_probe_fp64_tflops": (C.c_int, [_ctx, C.POINTER(C.c_double)]),
    "fsae_pool_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int]),
    "fsae_pool_destroy": (C.c_int, [C.c_void_p]),
    "fsae_pool_size": (C.c_int, [C.c_void_p]),
    "fsae_pool_ctx": (C.c_void_p, [C.c_void_p, C.c_int]),
    "fsae_pool_last_error": (C.c_char_p, [C.c_void_p]),
    "fsae_pool_set_track": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp, C.c_int, C.c_double]),
    "fsae_pool_set_params": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Params)]),
    "fsae_shard_range": (None, [C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "fsae_ltvmpc_host_pool": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, _ip, _ip,
                                        _dp, _dp, _dp, _dp, _dp, _dp, _ip, _dp, _dp, _ip, _bp, _bp]),
    "fsae_set_host_staging": (C.c_int, [_ctx, C.c_int]),
    "fsae_last_host_path": (C.c_int, [_ctx]),
}

_libs = {}


def load(path=None):
    """Load libfsae_mpc_b200.so (build it with `python -m fsae_mpc_b200.build`).  `path` (or the environment
    variable FSAE_LIB) selects another build of the same C-ABI, e.g. the cross-check library
    libfsae_mpc_b200_xcheck.so that the tests use for fsae_debug_set_kernel_version."""
    path = path or os.environ.get("FSAE_LIB") or LIB_PATH
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: the CUDA extension was not built "
            "(python -m fsae_mpc_b200.build).  fsae_mpc_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _libs[path] = lib
    return lib
