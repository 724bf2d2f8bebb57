"""Golden vectors for the CLOSED LOOP and the plant side, produced by the reference's own source:

  * main.m:24-58, :62-75, :84-88 (set-up) and the simulation loop main.m:91-190 are executed
    LITERALLY by oracle/mlab (`run_script_lines`), for MODEL = "KINEMATIC" and "DYNAMIC", with
    VISUALISE = false and N_simulation shortened (kinematic: without main.m:133, see run_main); the only substituted function is the third-party
    qpOASES MEX (oracle QP, as in make_reference_fixtures.py).  Everything else the loop calls --
    cartesian_to_curvilinear.m, closest_point.m, the speed ramp main.m:106-114, ltvmpc_*_curvilinear.m,
    pid_controller.m, integrate_cart_dyn.m, f_cart_dyn.m -- is the reference's file.
  * the plant functions on their own on seeded random inputs (incl. PID saturation, angle wrap).

    PYTHONPATH=. python scripts/make_reference_closed_loop.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from make_reference_fixtures import make_interp, GOLD, R  # noqa: E402

MAIN = R + "/main.m"


def run_main(model, track, n_sim, tracks):
    captured = {}
    ml = make_interp(captured)
    ml.paths += [R + "/vehicle_models/cartesian_dynamic", R + "/vehicle_models/cartesian_kinematic"]
    xs, ys, dl, L = tracks[track]
    env = {"x_spline": xs, "y_spline": ys, "dl": np.array([[dl]]), "L": np.array([[L]])}
    env["kappa"] = lambda s_, nargout=1: [ml.call("interpolate_curvature", s_, xs, ys, dl)]     # main.m:18
    ml.run_script_lines(MAIN, 25, 26, env)          # MODE, MODEL
    env["MODEL"] = model
    env["VISUALISE"] = np.array([[False]])
    ml.run_script_lines(MAIN, 29, 58, env)          # horizon, reference, initial guess
    ml.run_script_lines(MAIN, 62, 75, env)          # simulation state, preallocation
    ml.run_script_lines(MAIN, 84, 88, env)          # actuator PIDs
    env["N_simulation"] = np.array([[float(n_sim)]])
    if model == "DYNAMIC":
        ml.run_script_lines(MAIN, 91, 190, env)     # the loop, literally
    else:
        # main.m:133 reads slack_opt(4), which exists only for the dynamic model (N_soft = 4): with
        # MODEL = "KINEMATIC" the reference's own script stops there with an index error.  The loop body
        # is therefore executed without that one bookkeeping line (and without the NMPC branch :135-160 around
        # it, which MODE = "LTV-MPC" never takes): main.m:92-114, :120-132, :162-189 per i.
        for i in range(1, n_sim + 1):
            env["i"] = np.array([[float(i)]])
            ml.run_script_lines(MAIN, 92, 114, env)
            if float(np.asarray(env["s"]).ravel()[0]) >= L:      # the `break` of main.m:102-104
                break
            ml.run_script_lines(MAIN, 120, 132, env)
            ml.run_script_lines(MAIN, 162, 189, env)
    steps = int(np.asarray(env["i"]).ravel()[0])
    out = dict(n_list=env["n_list"].ravel()[:steps], exit_status=env["exit_status"].ravel()[:steps],
               objective=env["objective"].ravel()[:steps], x_history=env["x_history"][:steps],
               u_opt_history=env["u_opt_history"][:steps], x_opt_history=env["x_opt_history"][:steps],
               slack_n=env["slack_n"].ravel()[:steps], steps=steps,
               x_final=np.asarray(env["x"]).ravel(), x_opt_final=np.asarray(env["x_opt"]).ravel(),
               u_opt_final=np.asarray(env["u_opt"]).ravel())
    print(model, track, "steps", steps, "max|n|", np.abs(out["n_list"]).max(), "exit!=0:", int((out["exit_status"] != 0).sum()))
    return out, sorted(ml.calls)


def run_plant(tracks):
    ml = make_interp({})
    ml.paths += [R + "/vehicle_models/cartesian_dynamic"]
    rng = np.random.default_rng(11)
    out = {}
    # f_cart_dyn / integrate_cart_dyn
    X = np.column_stack([rng.uniform(-50, 50, 48), rng.uniform(-50, 50, 48), rng.uniform(-4, 4, 48),
                         rng.uniform(0.0, 25, 48), rng.uniform(-1.5, 1.5, 48), rng.uniform(-1.2, 1.2, 48),
                         rng.uniform(-0.4, 0.4, 48)])
    X[:4, 3:6] = 0.0                                   # standstill (the 0.01 regularisation of the slip angles)
    U = np.column_stack([rng.uniform(-2800, 2800, 48), rng.uniform(-0.8, 0.8, 48)])
    out["plant_x"], out["plant_u"] = X, U
    out["f_cart_dyn"] = np.stack([ml.call("f_cart_dyn", x.reshape(-1, 1), u.reshape(-1, 1)).ravel() for x, u in zip(X, U)])
    out["integrate_dt"] = np.array([0.005, 0.05])
    out["integrate_cart_dyn"] = np.stack([[ml.call("integrate_cart_dyn", x.reshape(-1, 1), u.reshape(-1, 1), h).ravel()
                                           for x, u in zip(X, U)] for h in out["integrate_dt"]])
    # pid_controller: both actuator settings of main.m:84-88, chained calls (status carried over)
    pid_in, pid_out = [], []
    for settings in ([16000.0, 0.0, 0.0, 2800.0], [80.0, 0.0, 0.0, 0.8], [3.0, 0.5, 0.2, 4.0]):
        status = [np.array([[0.0]]), np.array([[0.0]])]
        for _ in range(8):
            tgt, cur = rng.uniform(-2, 22), rng.uniform(-2, 22)
            o, status = ml.call("pid_controller", tgt, cur, [np.array([[s]]) for s in settings], status, nargout=2)
            pid_in.append(settings + [tgt, cur])
            pid_out.append([float(np.asarray(o).ravel()[0]), float(np.asarray(status[0]).ravel()[0]),
                            float(np.asarray(status[1]).ravel()[0])])
    out["pid_in"], out["pid_out"] = np.array(pid_in), np.array(pid_out)
    # cartesian_to_curvilinear / closest_point: points up to 1.5 m off the centre line, yaw up to several turns
    # away from the tangent (angdiff wrap), initial guesses up to 3 m off along the track
    from oracle import spline as sp, vehicle as vm      # only to PLACE the query points on the track
    c_in, c_out = [], []
    for ti, (name, (xs, ys, dl, L)) in enumerate(tracks.items()):
        for _ in range(24):
            s = rng.uniform(2.0, L - 2.0)
            n = rng.uniform(-1.5, 1.5)
            mu = rng.uniform(-0.5, 0.5) + rng.integers(-2, 3) * 2 * np.pi
            px, py, th = vm.curvilinear_to_cartesian(s, n, mu, xs, ys, dl)
            s0 = s + rng.uniform(-3, 3)
            r = ml.call("cartesian_to_curvilinear", float(px[0]), float(py[0]), float(th[0]), xs, ys, dl, s0, nargout=3)
            c_in.append([ti, float(px[0]), float(py[0]), float(th[0]), s0])
            c_out.append([float(np.asarray(v).ravel()[0]) for v in r])
    out["c2c_in"], out["c2c_out"] = np.array(c_in), np.array(c_out)
    print("plant functions:", ", ".join(sorted(ml.calls)))
    return out, sorted(ml.calls)


if __name__ == "__main__":
    t = dict(np.load(os.path.join(GOLD, "tracks.npz")))
    tracks = {n: (t[n + "_x"], t[n + "_y"], float(t[n + "_dl"]), float(t[n + "_L"])) for n in ("fsg2019", "fss2019", "fso2020")}
    res, executed = {}, set()
    o, ex = run_plant(tracks)
    executed.update(ex)
    res.update(o)
    for model, track, n_sim in (("KINEMATIC", "fsg2019", 60), ("DYNAMIC", "fss2019", 40)):
        o, ex = run_main(model, track, n_sim, tracks)
        executed.update(ex)
        for k, v in o.items():
            res[f"main_{model}_{k}"] = v
        res[f"main_{model}_track"] = np.array(track)
    res["executed"] = np.array(sorted(executed))
    np.savez_compressed(os.path.join(GOLD, "reference_m_closed_loop.npz"), **res)
    print("reference functions executed:", ", ".join(sorted(executed)))
