"""Pivot-order sensitivity: solve one batch with two kernel variants and dump the problem on which the
answers differ most (checked on the CPU against the oracle by scripts/sensitivity_check.py).
    python scripts/sensitivity.py kinematic|dynamic KV_A KV_B [B]"""
import os, sys, numpy as np
sys.path.insert(0, '.')
import fsae_mpc_b200 as fm
from fsae_mpc_b200 import workload as wl
model, kva, kvb = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
B = int(sys.argv[4]) if len(sys.argv) > 4 else 32768
track = "fsg2019" if model == "kinematic" else "fss2019"
mpc = fm.FsaeMpc(0)
for tid, (n, t) in enumerate(wl.load_tracks().items()):
    mpc.set_track(tid, t[0], t[1], t[2])
x0, xr, xl, ul = wl.perturbed_batch(model, track, B, 0)
tid = np.full(B, list(wl.load_tracks()).index(track), np.int32)
mpc.set_params(1, fm.default_params(fm.DYNAMIC if model == "dynamic" else fm.KINEMATIC))
pid = np.ones(B, np.int32)
step = mpc.ltvmpc_dynamic_curvilinear if model == "dynamic" else mpc.ltvmpc_kinetmatic_curvilinear
res = {}
for kv in (kva, kvb):
    mpc.set_kernel_version(kv)
    res[kv] = step(x0, xr, 0.05, xl, ul, track_id=tid, param_id=pid)
a, b = res[kva], res[kvb]
du = np.abs(a.u_opt - b.u_opt).reshape(B, -1).max(1)
order = np.argsort(-du)[:5]
print(f"worst |du| between kernel variants {kva} and {kvb}:", du[order], "at", order, " #(|du|>1e-6):", int((du > 1e-6).sum()))
print("fval", a.fval[order], b.fval[order], "iters", a.iters[order], b.iters[order])
i = int(order[0])
os.makedirs("gpurun_out", exist_ok=True)
np.savez("gpurun_out/worst.npz", model=model, track=track, x0=x0[i], x_ref=xr[i], x_lin=xl[i], u_lin=ul[i], u_a=a.u_opt[i], u_b=b.u_opt[i],
         f_a=a.fval[i], f_b=b.fval[i], s_a=a.slack_opt[i], s_b=b.slack_opt[i], du=du)
