"""Offline experiment (numpy, CPU): iteration counts of the dual active-set method on bench-like problems under
different pivot rules (which violated constraint enters).  Not product code."""
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
from oracle import spline as sp, ltv, qp as oqp
from fsae_mpc_b200 import workload as wl
from proto_gi_kform import gi_kform
model = sys.argv[1] if len(sys.argv) > 1 else "kinematic"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
trk = "fsg2019" if model == "kinematic" else "fss2019"
t = wl.load_tracks()[trk]
track = sp.Track(t[0], t[1], t[2], t[3])
x0, xr, xl, ul = wl.perturbed_batch(model, trk, B, 1000)
build = ltv.build_kinematic_qp if model == "kinematic" else ltv.build_dynamic_qp
rules = ["raw", "eucl", "norm"]
tot = {r: [0, 0, 0] for r in rules}
for b in range(B):
    Q = build(x0[b], xr[b].T, track.kappa, 0.05, xl[b].T, ul[b].T)
    n = Q['f'].size; ns = Q['n_soft']
    C, lo, up = oqp._stack(n, Q['xA'], Q['lb'], Q['ub'], Q['lbA'], Q['ubA'])
    W0 = [(n - ns + k, -1) for k in range(ns)]
    ref = None
    for r in rules:
        x, act, lam, it, st, (na, nd) = gi_kform(Q['H'], Q['f'], C, lo, up, W0, rule=r)
        if ref is None: ref = x
        err = np.max(np.abs(x - ref)[:n - ns])
        tot[r][0] += na; tot[r][1] += nd; tot[r][2] = max(tot[r][2], err)
        if st != 0: print("problem", b, "rule", r, "status", st)
for r in rules:
    print(f"{r:6s} adds/QP {tot[r][0]/B:.1f} drops/QP {tot[r][1]/B:.1f} iters {sum(tot[r][:2])/B:.1f}  max |dx| vs raw {tot[r][2]:.2e}")
