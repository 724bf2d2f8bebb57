"""Build the in-tree CUDA shared libraries (sm_100a only).

    python -m fsae_mpc_b200.build [--force] [-v]

  libfsae_mpc_b200.so         the product: host glue (capi.cu) + one translation unit per fused kernel family
  libfsae_mpc_b200_xcheck.so  the same plus the cross-check kernels (-DFSAE_XCHECK: the shared-memory operator
                              kernel fused_v1.cuh and the warp-count / block-size variants); tests only

The translation units compile in parallel (objects under build/obj).  nvcc cross-compiles without a GPU.
The .so files are git-ignored but travel with gpurun.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(ROOT, "build", "obj")
LIB = os.path.join(HERE, "libfsae_mpc_b200.so")
LIB_XCHECK = os.path.join(HERE, "libfsae_mpc_b200_xcheck.so")

# translation unit -> compiled with -DFSAE_XCHECK for the cross-check library?
KERNEL_TUS = ["k_kin40", "k_kin20", "k_kin80", "k_dyn40", "k_dyn20", "k_dyn80"]
XCHECK_VARIANT_TUS = {"capi", "k_kin40", "k_kin20"}        # these differ between the two libraries
PRODUCT = ["capi"] + KERNEL_TUS
XCHECK = ["capi", "k_xcheck"] + KERNEL_TUS

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-diag-suppress", "128,39"]


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(ROOT, "include", "fsae_mpc_b200.h"))
    return hs


def _obj_name(tu, xcheck):
    return os.path.join(OBJ, tu + ("_x" if xcheck and (tu in XCHECK_VARIANT_TUS or tu == "k_xcheck") else "") + ".o")


def _compile(tu, xcheck, nvcc, verbose):
    src = os.path.join(CSRC, tu + ".cu")
    out = _obj_name(tu, xcheck)
    cmd = [nvcc] + NVCC_FLAGS + (["-DFSAE_XCHECK"] if out.endswith("_x.o") else []) + \
          (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", out, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return res.stderr


def build(force=False, verbose=False, xcheck=True):
    """Compile csrc/*.cu -> libfsae_mpc_b200.so (+ the cross-check library).  Returns the product library path."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = max(os.path.getmtime(h) for h in _headers())
    jobs = {}
    for lib_tus, xc in ((PRODUCT, False),) + (((XCHECK, True),) if xcheck else ()):
        for tu in lib_tus:
            if not os.path.exists(os.path.join(CSRC, tu + ".cu")):
                continue
            out = _obj_name(tu, xc)
            src_t = max(hdr_t, os.path.getmtime(os.path.join(CSRC, tu + ".cu")))
            if force or not os.path.exists(out) or os.path.getmtime(out) < src_t:
                jobs[out] = (tu, xc)
    logs = []
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            logs = list(ex.map(lambda j: _compile(j[0], j[1], nvcc, verbose), jobs.values()))
    for lib, tus, xc in ((LIB, PRODUCT, False),) + (((LIB_XCHECK, XCHECK, True),) if xcheck else ()):
        objs = [_obj_name(t, xc) for t in tus if os.path.exists(os.path.join(CSRC, t + ".cu"))]
        if force or not os.path.exists(lib) or any(os.path.getmtime(o) > os.path.getmtime(lib) for o in objs):
            cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
