"""Which dynamic-model problems depend on the pivot order?  Solves one batch with two block sizes of the
dual active-set core and dumps the problem on which the two answers differ most (analysed on the CPU
against the oracle: scripts/dyn_sensitivity_check.py).   python scripts/dyn_sensitivity.py [B]"""
import os, sys, numpy as np
sys.path.insert(0, '.')
import fsae_mpc_b200 as fm
from fsae_mpc_b200 import workload as wl
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
mpc = fm.FsaeMpc(0)
for tid, (n, t) in enumerate(wl.load_tracks().items()):
    mpc.set_track(tid, t[0], t[1], t[2])
x0, xr, xl, ul = wl.perturbed_batch("dynamic", "fss2019", B, 0)
tid = np.full(B, list(wl.load_tracks()).index("fss2019"), np.int32)
mpc.set_params(1, fm.default_params(fm.DYNAMIC))      # parameter set 0 holds the kinematic defaults
pid = np.ones(B, np.int32)
res = {}
for kv in (21, 24):
    mpc.set_kernel_version(kv)
    res[kv] = mpc.ltvmpc_dynamic_curvilinear(x0, xr, 0.05, xl, ul, track_id=tid, param_id=pid)
a, b = res[21], res[24]
du = np.abs(a.u_opt - b.u_opt).reshape(B, -1).max(1)
order = np.argsort(-du)[:5]
print("worst |du| between block sizes 1 and 4:", du[order], "at", order, " #(|du|>1e-6):", int((du > 1e-6).sum()))
print("fval", a.fval[order], b.fval[order], "iters", a.iters[order], b.iters[order])
i = int(order[0])
os.makedirs("gpurun_out", exist_ok=True)
np.savez("gpurun_out/dyn_worst.npz", x0=x0[i], x_ref=xr[i], x_lin=xl[i], u_lin=ul[i], u_a=a.u_opt[i], u_b=b.u_opt[i],
         f_a=a.fval[i], f_b=b.fval[i], s_a=a.slack_opt[i], s_b=b.slack_opt[i], du=du)
