"""Generate tests/golden/*.npz from the ORACLE run in the build container.

Needs /root/reference/data/*.csv (track way-points are DATA of the reference, read here
once; only derived spline coefficients and oracle results are committed).  The GPU box has
no /root/reference: tests and bench read only the committed fixtures.

    PYTHONPATH=. python scripts/make_golden.py [--quick]
"""
import os
import sys
import time
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import spline as sp, closed_loop as cl, ltv  # noqa: E402

REF = "/root/reference/data"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
DT = 0.05


def tracks():
    d = {}
    for name in ("fsg2019", "fss2019", "fso2020"):
        tr = sp.Track.from_csv(f"{REF}/{name}.csv")
        d[name + "_x"] = tr.x_spline
        d[name + "_y"] = tr.y_spline
        d[name + "_dl"] = np.float64(tr.dl)
        d[name + "_L"] = np.float64(tr.L)
    np.savez_compressed(os.path.join(OUT, "tracks.npz"), **d)
    return d


def lap(model, name, every, n_stage, n_sim=1000):
    tr = sp.Track.from_csv(f"{REF}/{name}.csv")
    recs = []

    def rec(i, inp, out):
        sol = out[5]
        recs.append(dict(step=i, x0=inp["x0"], x_ref=inp["x_ref"], x_lin=inp["x_lin"], u_lin=inp["u_lin"],
                         u_opt=out[0], x_opt=out[1], exitflag=out[2], fval=out[3], slack=np.asarray(out[4]),
                         wsB=sol.workingSetB.astype(np.int8), wsC=sol.workingSetC.astype(np.int8),
                         iters=sol.iter))
    t = time.time()
    h = cl.run(tr, model, n_sim=n_sim, record=rec)
    print(f"{model} {name}: {h['steps']} steps in {time.time() - t:.1f}s, "
          f"exit!=0: {sum(1 for r in recs if r['exitflag'] != 0)}")
    sel = recs[::every]
    out = {k: np.stack([np.asarray(r[k]) for r in sel]) for k in sel[0]}
    out["lap_steps"] = np.int64(h["steps"])
    # stage outputs for a few problems (linearise / condense parity)
    build = ltv.build_kinematic_qp if model == "KINEMATIC" else ltv.build_dynamic_qp
    stage_idx = np.linspace(0, len(sel) - 1, n_stage).astype(int)
    st = {k: [] for k in ("A", "B", "d", "A_bar", "B_bar", "d_bar", "H", "f", "xA", "lbA", "ubA", "lb", "ub", "const")}
    for i in stage_idx:
        r = sel[i]
        q = build(r["x0"], r["x_ref"], tr.kappa, DT, r["x_lin"], r["u_lin"])
        for k in st:
            st[k].append(np.asarray(q[k]))
    for k in st:
        out["stage_" + k] = np.stack(st[k])
    out["stage_idx"] = stage_idx
    np.savez_compressed(os.path.join(OUT, f"{model.lower()}_lap_{name}.npz"), **out)
    return sel, tr


def perturbed(model, name, sel, tr, n, seed):
    """The bench workload in miniature: lap samples with perturbed initial states."""
    rng = np.random.default_rng(seed)
    step = ltv.ltvmpc_kinetmatic_curvilinear if model == "KINEMATIC" else ltv.ltvmpc_dynamic_curvilinear
    recs = []
    for j in range(n):
        r = sel[rng.integers(len(sel))]
        x0 = r["x0"].copy()
        x0[1] += rng.uniform(-0.3, 0.3)
        x0[2] += rng.uniform(-0.08, 0.08)
        x0[3] = max(0.5, x0[3] + rng.uniform(-1.5, 1.5))
        x0[-1] += rng.uniform(-0.05, 0.05)
        out = step(x0, r["x_ref"], tr.kappa, DT, r["x_lin"], r["u_lin"])
        sol = out[5]
        recs.append(dict(x0=x0, x_ref=r["x_ref"], x_lin=r["x_lin"], u_lin=r["u_lin"], u_opt=out[0],
                         x_opt=out[1], exitflag=out[2], fval=out[3], slack=np.asarray(out[4]),
                         wsB=sol.workingSetB.astype(np.int8), wsC=sol.workingSetC.astype(np.int8),
                         iters=sol.iter))
    out = {k: np.stack([np.asarray(r[k]) for r in recs]) for k in recs[0]}
    print(f"{model} perturbed: exit!=0: {(out['exitflag'] != 0).sum()} slack>0: {(out['slack'].max(1) > 0).sum()}")
    np.savez_compressed(os.path.join(OUT, f"{model.lower()}_perturbed_{name}.npz"), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    quick = "--quick" in sys.argv
    tracks()
    sel, tr = lap("KINEMATIC", "fsg2019", 8, 6, n_sim=80 if quick else 1000)
    perturbed("KINEMATIC", "fsg2019", sel, tr, 16 if quick else 96, 1)
    sel, tr = lap("DYNAMIC", "fss2019", 8, 3, n_sim=40 if quick else 1000)
    perturbed("DYNAMIC", "fss2019", sel, tr, 8 if quick else 48, 2)
    sel, tr = lap("DYNAMIC", "fsg2019", 16, 2, n_sim=40 if quick else 1000)
    sel, tr = lap("KINEMATIC", "fso2020", 16, 2, n_sim=40 if quick else 1000)
    sel, tr = lap("KINEMATIC", "fss2019", 16, 2, n_sim=40 if quick else 1000)
