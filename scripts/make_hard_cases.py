"""Regression fixtures tests/golden/hard_cases.npz: problems of the synthetic bench batches (workload.perturbed_batch,
seed 0) on which an active-set code is most likely to go wrong -- the longest pivot sequences (hundreds of
iterations, many partial steps) and the two problems on which earlier kernel variants stopped at a point whose
recomputed multiplier was negative (|du| ~ 1e-2).  Inputs were picked on the GPU (iteration counts) and are kept
under tests/golden/src/; the expected outputs are the ORACLE's.      python scripts/make_hard_cases.py"""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fsae_mpc_b200 import workload as wl
from oracle import spline as sp, ltv
G = os.path.join(ROOT, "tests", "golden")
tracks = {n: sp.Track(t[0], t[1], t[2], t[3]) for n, t in wl.load_tracks().items()}
out = {}
def run(tag, model, track, x0, xr, xl, ul):
    fn = ltv.ltvmpc_dynamic_curvilinear if model == "dynamic" else ltv.ltvmpc_kinetmatic_curvilinear
    U, X, F, S, E, IT = [], [], [], [], [], []
    for b in range(x0.shape[0]):
        u, x, ef, fv, sl, sol = fn(x0[b], xr[b].T, tracks[track].kappa, 0.05, xl[b].T, ul[b].T)
        U.append(u); X.append(x); F.append(fv); S.append(sl); E.append(ef); IT.append(sol.iter)
        print(tag, b, "exitflag", ef, "oracle iterations", sol.iter, "kkt", {k: float(v) for k, v in sol.kkt.items()})
    out.update({f"{tag}_x0": x0, f"{tag}_x_ref": xr, f"{tag}_x_lin": xl, f"{tag}_u_lin": ul, f"{tag}_u_opt": np.array(U),
                f"{tag}_x_opt": np.array(X), f"{tag}_fval": np.array(F), f"{tag}_slack": np.array(S), f"{tag}_exitflag": np.array(E)})
d = np.load(os.path.join(G, "src", "hard_dyn_inputs.npz"))
run("dyn", "dynamic", "fss2019", d["x0"], d["x_ref"], d["x_lin"], d["u_lin"])
d = np.load(os.path.join(G, "src", "hard_kin_inputs.npz"))
p = np.load(os.path.join(G, "src", "hard_kin_pivot_inputs.npz"))
cat = lambda k: np.concatenate([d[k], p[k][None]])
run("kin", "kinematic", "fsg2019", cat("x0"), cat("x_ref"), cat("x_lin"), cat("u_lin"))
np.savez_compressed(os.path.join(G, "hard_cases.npz"), **out)
