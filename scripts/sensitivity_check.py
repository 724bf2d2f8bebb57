"""CPU side of scripts/sensitivity.py: which of the two dumped answers is the oracle's?"""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fsae_mpc_b200 import workload as wl
from oracle import spline as sp, ltv
d = np.load(os.path.join(ROOT, "gpurun_out", "worst.npz"))
model, track = str(d["model"]), str(d["track"])
t = wl.load_tracks()[track]
trk = sp.Track(t[0], t[1], t[2], t[3])
fn = ltv.ltvmpc_dynamic_curvilinear if model == "dynamic" else ltv.ltvmpc_kinetmatic_curvilinear
u, x, ef, fv, sl, sol = fn(d["x0"], d["x_ref"].T, trk.kappa, 0.05, d["x_lin"].T, d["u_lin"].T)
print("problems with |du| > 1e-6:", int((d["du"] > 1e-6).sum()), "max", d["du"].max())
print("oracle exitflag", ef, "fval", fv, "iters", sol.iter)
for n in "ab":
    print(f"variant {n}: |u - u_oracle| {np.abs(d['u_' + n].reshape(-1) - u).max():.3e}  fval - fval_oracle {float(d['f_' + n]) - fv:.3e}  slack {d['s_' + n]}")
