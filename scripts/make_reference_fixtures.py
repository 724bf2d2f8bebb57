"""Golden vectors produced by the REFERENCE'S OWN SOURCE: runs the unmodified .m files under
/root/reference through oracle/mlab (a MATLAB-subset interpreter; the image has no MATLAB) and
stores what they compute.  The only substituted function is the third-party `qpOASES` MEX
(Windows binary): the reference's call at ltvmpc_*_curvilinear.m:52 is intercepted, its seven
arguments (H, f, xA, lb, ub, lbA, ubA) are recorded, and the oracle QP solver supplies the
minimiser so the reference's post-processing (x_opt, fval + const, slack_opt) also runs.

    PYTHONPATH=. python scripts/make_reference_fixtures.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.mlab.interp import Matlab  # noqa: E402
from oracle import qp as oqp  # noqa: E402

R = "/root/reference"
PATHS = [R + "/spline", R + "/vehicle_models/curvilinear_kinematic", R + "/vehicle_models/curvilinear_dynamic",
         R + "/mpc/ltv", R + "/mpc/ltv/kinematic", R + "/mpc/ltv/dynamic", R + "/util", R + "/vehicle_models"]
GOLD = os.path.join(ROOT, "tests", "golden")
DT = 0.05


def make_interp(captured):
    def qpoases_override(H, g, A, lb, ub, lbA, ubA, nargout=1):
        captured.update(H=H, f=g.ravel(), xA=A, lb=lb.ravel(), ub=ub.ravel(), lbA=lbA.ravel(), ubA=ubA.ravel())
        sol = oqp.qpoases(H, g.ravel(), A, lb.ravel(), ub.ravel(), lbA.ravel(), ubA.ravel())
        outs = [sol.x.reshape(-1, 1), np.array([[sol.fval]]), np.array([[float(sol.exitflag)]]),
                np.array([[float(sol.iter)]]), sol.lam.reshape(-1, 1), 0.0]
        return outs[:max(nargout, 1)]
    return Matlab(PATHS, overrides={"qpOASES": qpoases_override})


def run(model, track, fixture, picks):
    t = dict(np.load(os.path.join(GOLD, "tracks.npz")))
    xs, ys, dl = t[track + "_x"], t[track + "_y"], float(t[track + "_dl"])
    g = dict(np.load(os.path.join(GOLD, fixture)))
    captured = {}
    ml = make_interp(captured)
    kappa = lambda s_, nargout=1: [ml.call("interpolate_curvature", s_, xs, ys, dl)]  # main.m:18
    lin = "rk2_kinematic_curvilinear" if model == "kinematic" else "rk4_dynamic_curvilinear"
    step = "ltvmpc_kinetmatic_curvilinear" if model == "kinematic" else "ltvmpc_dynamic_curvilinear"
    rec = {k: [] for k in ("x0", "x_ref", "x_lin", "u_lin", "A", "B", "d", "A_bar", "B_bar", "d_bar", "H", "f", "xA",
                           "lb", "ub", "lbA", "ubA", "u_opt", "x_opt", "fval", "slack", "exitflag")}
    for b in picks:
        x0, xr, xl, ul = g["x0"][b], g["x_ref"][b], g["x_lin"][b], g["u_lin"][b]
        A, B, d = ml.call(lin, xl, ul, kappa, DT, nargout=3)
        Ab, Bb, db = ml.call("sequential_integration", A, B, d, DT, nargout=3)
        u_opt, x_opt, _, ef, fval, slack = ml.call(step, x0.reshape(-1, 1), xr, kappa, DT, xl, ul, 0.0, nargout=6)
        for k, v in dict(x0=x0, x_ref=xr, x_lin=xl, u_lin=ul, A=A, B=B, d=d, A_bar=Ab, B_bar=Bb, d_bar=db.ravel(),
                         u_opt=u_opt.ravel(), x_opt=x_opt.ravel(), fval=float(fval.ravel()[0]), slack=slack.ravel(),
                         exitflag=float(np.asarray(ef).ravel()[0]), **captured).items():
            rec[k].append(np.asarray(v, dtype=np.float64))
        print(model, track, "problem", b, "exitflag", rec["exitflag"][-1])
    out = {k: np.stack(v) for k, v in rec.items()}
    out["executed"] = np.array(sorted(ml.calls))
    np.savez_compressed(os.path.join(GOLD, f"reference_m_{model}_{track}.npz"), **out)
    print("reference functions executed:", ", ".join(sorted(ml.calls)))




def run_obtain_reference():
    """util/obtain_reference.m executed on synthetic plans (the reference's own plan comes from its IPOPT
    minimum-time planner, a Windows MEX binary): smooth lap-like state samples, segment times between a
    fraction of dt and several dt, starting points before, inside and beyond one lap."""
    ml = Matlab([R + "/util"])
    rng = np.random.default_rng(5)
    cases = []
    for N_s, ds, N_t, dt in ((50, 2.5, 40, 0.05), (300, 1.0, 40, 0.05), (128, 0.4, 80, 0.05), (64, 3.0, 20, 0.1)):
        ph = 2 * np.pi * np.arange(N_s) / N_s
        plan = np.zeros((8, N_s))
        for k in range(8):
            plan[k] = rng.normal() * np.sin((k % 3 + 1) * ph + rng.uniform(0, 6)) + 0.1 * rng.normal(size=N_s)
        plan[2] = 12.0 + 6.0 * np.sin(ph)                       # x_d > 0
        t = ds / plan[2] * (1.0 + 0.3 * rng.uniform(-1, 1, N_s))
        x = plan.reshape(-1, order="F").reshape(-1, 1)
        for s0 in (0.0, 0.37 * ds, 7.9 * ds, ds * N_s - 0.2 * ds, 1.6 * ds * N_s, float(rng.uniform(0, ds * N_s))):
            out = ml.call("obtain_reference", x, ds, float(N_s), t.reshape(-1, 1), s0, dt, float(N_t))
            cases.append(dict(x=x.ravel(), t=t, ds=ds, N_s=N_s, s0=s0, dt=dt, N_t=N_t, x_ref=np.asarray(out)))
    np.savez_compressed(os.path.join(GOLD, "reference_m_obtain_reference.npz"), n=len(cases),
                        **{f"c{i}_{k}": np.asarray(v) for i, c in enumerate(cases) for k, v in c.items()})
    print("obtain_reference:", len(cases), "cases")


if __name__ == "__main__":
    if "--only-reference" not in sys.argv:
        run("kinematic", "fsg2019", "kinematic_lap_fsg2019.npz", [3, 20, 41, 60, 77])
        run("dynamic", "fss2019", "dynamic_lap_fss2019.npz", [5, 23, 40])
    run_obtain_reference()
