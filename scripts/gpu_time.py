import sys, time, numpy as np
sys.path.insert(0, '.')
import fsae_mpc_b200 as fm
from fsae_mpc_b200 import workload as wl
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
model = sys.argv[2] if len(sys.argv) > 2 else "kinematic"
track = sys.argv[3] if len(sys.argv) > 3 else "fsg2019"
mpc = fm.FsaeMpc(0)
for tid, (n, t) in enumerate(wl.load_tracks().items()):
    mpc.set_track(tid, t[0], t[1], t[2])
x0, xr, xl, ul = wl.perturbed_batch(model, track, B, 0)
tid = np.full(B, list(wl.load_tracks()).index(track), np.int32)
pid = None
if model == "dynamic":
    mpc.set_params(1, fm.default_params(fm.DYNAMIC)); pid = np.ones(B, np.int32)
step = mpc.ltvmpc_kinetmatic_curvilinear if model == "kinematic" else mpc.ltvmpc_dynamic_curvilinear
import os
if os.environ.get("FSAE_KV"): mpc.set_kernel_version(int(os.environ["FSAE_KV"]))
for rep in range(4):
    mpc.counters(reset=True)
    t = time.time()
    r = step(x0, xr, 0.05, xl, ul, track_id=tid, param_id=pid)
    dt_host = time.time() - t
    ms = mpc.last_kernel_ms
    a, d, rf = mpc.counters()
    print(f"B={B} last chunk of the pipelined host call: kernel {ms:.2f} ms (see gpu_time_dev.py for device-resident timing); host call {dt_host*1e3:.1f} ms; exit!=0 {(r.exitflag!=0).sum()} iters mean {r.iters.mean():.1f} max {r.iters.max()} adds/QP {a/B:.1f} drops/QP {d/B:.1f} refresh/QP {rf/B:.2f} slack>0 {(r.slack_opt>0).sum()}")
