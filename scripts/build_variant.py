"""Build a VARIANT of the product library with extra -D flags (A/B measurements of kernel options):
    python scripts/build_variant.py noreuse -DFSAE_REUSE=0     ->  build/libfsae_noreuse.so
Run a script against it with  FSAE_LIB=build/libfsae_noreuse.so python ...   (the product .so is not touched)."""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fsae_mpc_b200 import build as b
name, extra = sys.argv[1], sys.argv[2:]
only = [a[len("--only="):].split(",") for a in extra if a.startswith("--only=")]
extra = [a for a in extra if not a.startswith("--only=")]
obj = os.path.join(ROOT, "build", "variants", name)
os.makedirs(obj, exist_ok=True)
tus = [t for t in b.PRODUCT if os.path.exists(os.path.join(b.CSRC, t + ".cu"))]
def cc(t):
    src_o = os.path.join(obj, t + ".o")
    if only and t not in only[0] and os.path.exists(b._obj_name(t, False)):
        return b._obj_name(t, False)                  # reuse the product object for TUs the flag does not touch
    subprocess.run(["nvcc"] + b.NVCC_FLAGS + extra + ["-c", "-o", src_o, os.path.join(b.CSRC, t + ".cu")], check=True)
    return src_o
with ThreadPoolExecutor(8) as ex:
    objs = list(ex.map(cc, tus))
out = os.path.join(ROOT, "build", f"libfsae_{name}.so")
subprocess.run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs, check=True)
print(out)
