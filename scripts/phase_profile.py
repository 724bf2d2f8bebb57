"""Per-phase cycle breakdown of the dual active-set loop (warp 0's clock64 deltas), from a
separate -DFSAE_PROFILE build (the product .so is not touched).
    python scripts/phase_profile.py [B]"""
import ctypes as C, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
prof_so = os.path.join(ROOT, "build", "libfsae_prof.so")
CSRC = os.path.join(ROOT, "fsae_mpc_b200", "csrc")
if not (os.environ.get("FSAE_PROF_NOBUILD") and os.path.exists(prof_so)) and (not os.path.exists(prof_so) or os.path.getmtime(prof_so) < max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC))):
    # the phase counters are __device__ globals: the profile build is ONE translation unit (all TUs included)
    os.makedirs(os.path.dirname(prof_so), exist_ok=True)
    unity = os.path.join(ROOT, "build", "prof_unity.cu")
    # only the kernel families being profiled are compiled (FSAE_PROF_TUS, default kin40 + dyn40); the others are stubs
    want = os.environ.get("FSAE_PROF_TUS", "k_kin40,k_dyn40").split(",")
    allk = {"k_kin40": "launch_kin40", "k_kin20": "launch_kin20", "k_kin80": "launch_kin80", "k_dyn40": "launch_dyn40",
            "k_dyn20": "launch_dyn20", "k_dyn80": "launch_dyn80"}
    with open(unity, "w") as fh:
        fh.write(f'#include "{os.path.join(CSRC, "capi.cu")}"\n#include "{os.path.join(CSRC, "k_xcheck.cu")}"\n')
        for k, fn in allk.items():
            if k in want:
                fh.write(f'#include "{os.path.join(CSRC, k + ".cu")}"\n')
            else:
                fh.write(f'namespace fsae {{ cudaError_t {fn}(const BatchArgs&, cudaStream_t, int) {{ return cudaErrorNotSupported; }} }}\n')
        for k in ("kin80", "dyn80"):
            if "k_" + k not in want:
                fh.write(f'namespace fsae {{ size_t slab_{k}() {{ return 1; }} }}\n')
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-DFSAE_PROFILE"] +
                   (["-DFSAE_XCHECK"] if os.environ.get("FSAE_PROF_XCHECK") else []) + ["-lineinfo",
                    "-diag-suppress", "128,39", "-shared", "-Xcompiler", "-fPIC", "-o", prof_so, unity], check=True)
os.environ["FSAE_LIB"] = prof_so
import fsae_mpc_b200 as fm
from fsae_mpc_b200 import workload as wl
B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
mpc = fm.FsaeMpc(0)
if os.environ.get("FSAE_KV"): mpc.set_kernel_version(int(os.environ["FSAE_KV"]))
for tid, (n, t) in enumerate(wl.load_tracks().items()):
    mpc.set_track(tid, t[0], t[1], t[2])
lib = mpc._lib
lib.fsae_profile_read.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_int]
lib.fsae_profile_read_stages.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_int]
sg = (C.c_uint64 * 16)()
x0, xr, xl, ul = wl.perturbed_batch("kinematic", "fsg2019", B, 0)
if os.environ.get("FSAE_N") == "80":          # horizon 80: the committed fsg2019 lap, unperturbed
    g = dict(np.load(os.path.join(wl.GOLDEN, "kinematic_lap_fsg2019_N80.npz")))
    pick = np.arange(B) % g["x0"].shape[0]
    tr_ = lambda a: np.ascontiguousarray(a.transpose(0, 2, 1)[pick])
    x0, xr, xl, ul = g["x0"][pick].copy(), tr_(g["x_ref"]), tr_(g["x_lin"]), tr_(g["u_lin"])
out = (C.c_uint64 * 16)()
step, kw = mpc.ltvmpc_kinetmatic_curvilinear, {}
if os.environ.get("FSAE_MODEL") == "dynamic":
    x0, xr, xl, ul = wl.perturbed_batch("dynamic", "fss2019", B, 0)
    mpc.set_params(1, fm.default_params(fm.DYNAMIC))
    step, kw = mpc.ltvmpc_dynamic_curvilinear, dict(track_id=np.ones(B, np.int32), param_id=np.ones(B, np.int32))
step(x0, xr, 0.05, xl, ul, **kw)
lib.fsae_profile_read(mpc._ctx, out, 1)
mpc.counters(reset=True)
lib.fsae_profile_read_stages(mpc._ctx, sg, 1)
r = step(x0, xr, 0.05, xl, ul, **kw)
lib.fsae_profile_read(mpc._ctx, out, 1)
adds, drops, refr = mpc.counters()
names = ["loop/refresh/end-of-block barrier", "P1 search (policy)", "P1 top-KB argmin+barrier", "P2 normals", "(unused)",
         "P3 y=M'n (+barrier)", "P4 step lengths", "P5 z=J2y2, x update", "P6a add update", "queue transform + advance", "P6b drop"]
tot = sum(out[i] for i in range(11))
print(f"searches/QP {out[12]/B:.1f}  piggy-backed adds/QP {out[13]/B:.1f}  piggy aborts (partial step)/QP {out[14]/B:.2f}  skipped (no longer violated)/QP {out[15]/B:.2f}")
it = r.iters.sum()
print(f"B={B}: iterations {it} (adds {adds}, drops {drops}), cycles in loop per QP {tot/B:.0f}, per iteration {tot/it:.0f}")
for i, n in enumerate(names):
    print(f"  {n:34s} {out[i]/it:8.0f} cyc/iter  {100*out[i]/tot:5.1f}%")
lib.fsae_profile_read_stages(mpc._ctx, sg, 1)
snames = ["load (TMA)", "linearise", "free response + B_bar chains", "g, bounds, row norms", "H build", "factor + layout", "initial point", "active-set loop", "outputs"]
stot = sum(sg[i] for i in range(9))
print(f"kernel stages, cycles per QP (tid 0): total {stot/B:.0f}")
for i, n in enumerate(snames):
    print(f"  {n:34s} {sg[i]/B:9.0f} cyc  {100*sg[i]/stot:5.1f}%")
print("recursion warp phases (cycles per QP): " + ", ".join(f"ph{i+1} {sg[9+i]/B:.0f}" for i in range(4)))
print("drop phases (cycles per QP): column+barrier %.0f, H k (symv) %.0f, M'(Hk) %.0f, update %.0f" % tuple(sg[12 + i] / B for i in range(4)))
