import sys, time, numpy as np
sys.path.insert(0, '.')
import fsae_mpc_b200 as fm
from fsae_mpc_b200 import workload as wl
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mpc = fm.FsaeMpc(0)
for tid, (n, t) in enumerate(wl.load_tracks().items()):
    mpc.set_track(tid, t[0], t[1], t[2])
x0, xr, xl, ul = wl.perturbed_batch("kinematic", "fsg2019", B, 0)
for rep in range(4):
    mpc.counters(reset=True)
    t = time.time()
    r = mpc.ltvmpc_kinetmatic_curvilinear(x0, xr, 0.05, xl, ul)
    dt_host = time.time() - t
    ms = mpc.last_kernel_ms
    a, d, rf = mpc.counters()
    print(f"B={B} kernel {ms:.2f} ms  -> {B/ms*1e3:.0f} QP/s ; host call {dt_host*1e3:.1f} ms; exit!=0 {(r.exitflag!=0).sum()} iters mean {r.iters.mean():.1f} max {r.iters.max()} adds/QP {a/B:.1f} drops/QP {d/B:.1f} refresh/QP {rf/B:.2f} slack>0 {(r.slack_opt>0).sum()}")
