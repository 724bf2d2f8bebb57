"""Wide set of golden vectors produced by the REFERENCE'S OWN SOURCE (round 2): the unmodified .m
files under /root/reference executed by oracle/mlab on

  * >= 50 problems per model at the default horizon (perturbed start states, the hard cases with
    the longest pivot sequences, laps of the other two tracks),
  * horizons 20 and 80 for both models,
  * all three linearisation schemes euler_/rk2_/rk4_*_curvilinear.m for both models.

Per problem the reference computes A, B, d (linearise), A_bar, B_bar, d_bar (sequential_integration.m),
H, f, xA, lb, ub, lbA, ubA (the seven arguments of the qpOASES call at ltvmpc_*_curvilinear.m:52,
intercepted) and its own post-processing of the minimiser (:57-60).  The two big matrices are
stored as PROBES  xA @ V  and  B_bar @ V  with a seeded dense V (nV x 3, stored in the file): any
wrong entry changes the probe, and the fixture stays a few MB.

    PYTHONPATH=. python scripts/make_reference_wide.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from make_reference_fixtures import make_interp, GOLD, DT  # noqa: E402

TRACKS = ("fsg2019", "fss2019", "fso2020")


def ramp_reference(x0, N_x, N):
    """main.m:107-114 (restated only to build INPUTS of the horizon-80 dynamic problems)."""
    x_ref = np.zeros((N_x, N))
    k = np.arange(1, N + 1)
    x_ref[3] = np.minimum(x0[3] + 10 * DT * k, 20.0) if x0[3] < 20.0 else np.maximum(x0[3] - 10 * DT * k, 20.0)
    x_ref[0] = x0[0] + np.cumsum(x_ref[3] * DT)
    return x_ref


def problems(model, N):
    """-> list of (track, x0, x_ref, x_lin, u_lin) in MATLAB shapes (NX,N)."""
    ld = lambda f: dict(np.load(os.path.join(GOLD, f)))
    out = []
    take = lambda g, idx, tr: [out.append((tr, g["x0"][i], g["x_ref"][i], g["x_lin"][i], g["u_lin"][i])) for i in idx]
    if model == "kinematic" and N == 40:
        take(ld("kinematic_perturbed_fsg2019.npz"), range(0, 96, 2)[:40], "fsg2019")
        h = ld("hard_cases.npz")
        for i in range(4):
            out.append(("fsg2019", h["kin_x0"][i], h["kin_x_ref"][i].T, h["kin_x_lin"][i].T, h["kin_u_lin"][i].T))
        g = ld("kinematic_lap_fso2020.npz"); take(g, np.linspace(1, g["x0"].shape[0] - 1, 8).astype(int), "fso2020")
        g = ld("kinematic_lap_fss2019.npz"); take(g, np.linspace(1, g["x0"].shape[0] - 1, 8).astype(int), "fss2019")
    elif model == "kinematic" and N == 20:
        g = ld("kinematic_lap_fsg2019_N20.npz"); take(g, np.linspace(0, g["x0"].shape[0] - 1, 8).astype(int), "fsg2019")
    elif model == "kinematic" and N == 80:
        g = ld("kinematic_lap_fsg2019_N80.npz"); take(g, [1, 5, 9, 12], "fsg2019")
    elif model == "dynamic" and N == 40:
        take(ld("dynamic_perturbed_fss2019.npz"), range(0, 48)[:36], "fss2019")
        h = ld("hard_cases.npz")
        for i in range(4):
            out.append(("fss2019", h["dyn_x0"][i], h["dyn_x_ref"][i].T, h["dyn_x_lin"][i].T, h["dyn_u_lin"][i].T))
        g = ld("dynamic_lap_fsg2019.npz"); take(g, np.linspace(1, g["x0"].shape[0] - 1, 12).astype(int), "fsg2019")
    elif model == "dynamic" and N == 20:
        g = ld("dynamic_lap_fss2019.npz")
        for i in np.linspace(2, g["x0"].shape[0] - 1, 8).astype(int):
            out.append(("fss2019", g["x0"][i], g["x_ref"][i][:, :20], g["x_lin"][i][:, :20], g["u_lin"][i][:, :20]))
    elif model == "dynamic" and N == 80:
        # horizon-80 inputs: the predictions of lap steps i and i+40 joined (a plausible linearisation
        # trajectory of 80 steps), the reference ramp of main.m over 80 steps
        g = ld("dynamic_lap_fss2019.npz")
        for i in (3, 9, 14):
            xl = np.concatenate([g["x_lin"][i], g["x_lin"][i + 40]], axis=1)
            ul = np.concatenate([g["u_lin"][i], g["u_lin"][i + 40]], axis=1)
            out.append(("fss2019", g["x0"][i], ramp_reference(g["x0"][i], 7, 80), xl, ul))
    return out


def run_group(model, N, tracks, V):
    captured = {}
    ml = make_interp(captured)
    step = "ltvmpc_kinetmatic_curvilinear" if model == "kinematic" else "ltvmpc_dynamic_curvilinear"
    lin = ("rk2_kinematic_curvilinear" if model == "kinematic" else "rk4_dynamic_curvilinear")
    rec = {}
    add = lambda k, v: rec.setdefault(k, []).append(np.asarray(v, dtype=np.float64))
    for b, (tr, x0, xr, xl, ul) in enumerate(problems(model, N)):
        xs, ys, dl = tracks[tr]
        kappa = lambda s_, nargout=1: [ml.call("interpolate_curvature", s_, xs, ys, dl)]      # main.m:18
        A, B, d = ml.call(lin, xl, ul, kappa, DT, nargout=3)
        Ab, Bb, db = ml.call("sequential_integration", A, B, d, DT, nargout=3)
        u_opt, x_opt, _, ef, fval, slack = ml.call(step, x0.reshape(-1, 1), xr, kappa, DT, xl, ul, 0.0, nargout=6)
        nU = Bb.shape[1]
        for k, v in dict(track=TRACKS.index(tr), x0=x0, x_ref=xr, x_lin=xl, u_lin=ul, A=A, B=B, d=d, A_bar=Ab,
                         d_bar=db.ravel(), B_bar_probe=Bb @ V[:nU], H=captured["H"], f=captured["f"],
                         xA_probe=captured["xA"] @ V, xA_abs_sum=np.abs(captured["xA"]).sum(axis=1),
                         lb=captured["lb"], ub=captured["ub"], lbA=captured["lbA"], ubA=captured["ubA"],
                         u_opt=u_opt.ravel(), x_opt=x_opt.ravel(), fval=float(fval.ravel()[0]), slack=slack.ravel(),
                         exitflag=float(np.asarray(ef).ravel()[0])).items():
            add(k, v)
        print(model, N, "problem", b, tr, "exitflag", rec["exitflag"][-1], flush=True)
    return {k: np.stack(v) for k, v in rec.items()}, sorted(ml.calls)


def run_schemes(model, tracks):
    """euler_/rk2_/rk4_*_curvilinear.m on 12 problems of the default horizon."""
    ml = make_interp({})
    short = "kinematic" if model == "kinematic" else "dynamic"
    pr = problems(model, 40)
    pick = list(range(0, len(pr), max(1, len(pr) // 12)))[:12]
    rec = {"idx": np.array(pick)}
    for sch in ("euler", "rk2", "rk4"):
        As, Bs, ds = [], [], []
        for b in pick:
            tr, x0, xr, xl, ul = pr[b]
            xs, ys, dl = tracks[tr]
            kappa = lambda s_, nargout=1: [ml.call("interpolate_curvature", s_, xs, ys, dl)]
            A, B, d = ml.call(f"{sch}_{short}_curvilinear", xl, ul, kappa, DT, nargout=3)
            As.append(A); Bs.append(B); ds.append(d)
        rec[f"{sch}_A"], rec[f"{sch}_B"], rec[f"{sch}_d"] = np.stack(As), np.stack(Bs), np.stack(ds)
        print(model, sch, "done", flush=True)
    return rec, sorted(ml.calls)


if __name__ == "__main__":
    t = dict(np.load(os.path.join(GOLD, "tracks.npz")))
    tracks = {n: (t[n + "_x"], t[n + "_y"], float(t[n + "_dl"])) for n in TRACKS}
    out, executed = {}, set()
    for model, nS in (("kinematic", 1), ("dynamic", 4)):
        for N in (40, 20, 80):
            V = np.random.default_rng(1000 + N + nS).standard_normal((2 * N + nS, 3))
            r, ex = run_group(model, N, tracks, V)
            executed.update(ex)
            out[f"{model}_N{N}_V"] = V
            for k, v in r.items():
                out[f"{model}_N{N}_{k}"] = v
        r, ex = run_schemes(model, tracks)
        executed.update(ex)
        for k, v in r.items():
            out[f"{model}_schemes_{k}"] = v
    out["executed"] = np.array(sorted(executed))
    np.savez_compressed(os.path.join(GOLD, "reference_m_wide.npz"), **out)
    print("reference functions executed:", ", ".join(sorted(executed)))
