"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/fsae_mpc_b200.h declares (no compute calls without a GPU), the ctypes Params struct
matches the C struct, the MEX gateway compiles against the header, and the product has no
CPU fallback / does not import the oracle."""
import ctypes as C
import os
import re
import shutil
import subprocess
import tempfile

import pytest

from conftest import ROOT

HDR = os.path.join(ROOT, "include", "fsae_mpc_b200.h")


def declared_symbols():
    txt = open(HDR).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fsae_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_every_declared_symbol():
    from fsae_mpc_b200 import build, _lib
    lib_path = build.build()
    assert os.path.exists(lib_path)
    lib = C.CDLL(lib_path)
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    # the ctypes table covers exactly the header
    assert sorted(_lib.SIGNATURES) == syms


def test_product_library_ships_only_product_kernels():
    """The shared-memory cross-check kernel (fused_v1.cuh) and the tuning variants live in the separate
    cross-check build; the product .so must not carry them."""
    from fsae_mpc_b200 import build
    build.build()
    prod = open(build.LIB, "rb").read()
    xchk = open(build.LIB_XCHECK, "rb").read()
    assert b"ltvmpc_fused_v1_kernel" not in prod and b"ltvmpc_fused_v1_kernel" in xchk
    assert prod.count(b"ltvmpc_fused_v2_kernel") > 0
    # kinematic N = 40 variants (8 / 4 warps, block adds) only in the cross-check build
    assert b"KinModelELi40ELi2ELi4ELi1" not in prod and b"KinModelELi40ELi2ELi4ELi1" in xchk


def test_params_struct_layout_matches_c():
    from fsae_mpc_b200 import _lib
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "fsae_mpc_b200.h"
    int main(void) {
        printf("%zu %zu %zu %zu %zu %zu\n", sizeof(fsae_params), offsetof(fsae_params, Q), offsetof(fsae_params, u_lb),
               offsetof(fsae_params, ay_max), offsetof(fsae_params, lin_scheme), offsetof(fsae_params, flat_eps));
        return 0;
    }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")], check=True)
        out = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout.split()
    P = _lib.Params
    assert [int(v) for v in out] == [C.sizeof(P), P.Q.offset, P.u_lb.offset, P.ay_max.offset,
                                     P.lin_scheme.offset, P.flat_eps.offset]


def test_default_params_are_the_reference_constants():
    import fsae_mpc_b200 as fm
    p = fm.default_params(fm.KINEMATIC)
    assert (p.lr, p.lf) == (0.6183, 0.8672)                       # f_curv_kin.m:13-14
    assert list(p.Q)[:5] == [5, 250, 2000, 0, 0]                  # ltvmpc_kinetmatic_curvilinear.m:32
    assert list(p.Q_terminal)[:3] == [50, 2500, 20000]            # :33
    assert list(p.R) == [10, 10] and p.R_soft[0] == 1e8           # :34-35
    assert (p.u_lb[0], p.u_ub[1], p.n_ub, p.ay_max) == (-10, 0.4, 0.75, 5.0)
    assert p.lin_scheme == fm.LIN_RK2
    d = fm.default_params(fm.DYNAMIC)
    assert list(d.R_soft) == [1e8, 1e6, 1e6, 1e4] and d.lin_scheme == fm.LIN_RK4   # ltvmpc_dynamic_curvilinear.m:35,38
    assert (d.mass, d.inertia, d.ac_max, d.al_max, d.slip_max) == (280.0, 200.0, 9.163, 10.0, 0.1)


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    import fsae_mpc_b200 as fm
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(fm.FsaeError):
        fm.FsaeMpc(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "fsae_mpc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "ltvmpc_oracle" not in txt, f


def test_mex_gateway_compiles_against_the_header():
    """matlab/fsae_mpc_b200_mex.c is the reference-side binding; no MATLAB here, so it is
    type-checked against a declarations-only mex.h stand-in."""
    with tempfile.TemporaryDirectory() as d:
        shutil.copy(os.path.join(ROOT, "matlab", "mex_stub.h"), os.path.join(d, "mex.h"))
        subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-I", d, "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "matlab", "fsae_mpc_b200_mex.c")], check=True)
