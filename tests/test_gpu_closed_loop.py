"""GPU: the batch driver (main.m's closed loop on device) against the oracle's main.m
restatement, vehicle by vehicle, from perturbed start states."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("model_name,n_sim", [("KINEMATIC", 70), ("DYNAMIC", 40)])
def test_closed_loop_batch_matches_oracle(mpc, fsg, model_name, n_sim):
    import fsae_mpc_b200 as fm
    from oracle import closed_loop as cl, vehicle as vm
    model = fm.KINEMATIC if model_name == "KINEMATIC" else fm.DYNAMIC
    pid = 6 + model
    mpc.set_params(pid, fm.default_params(model))
    # vehicle 0 is main.m's own start (x = zeros(7,1)); the others start displaced on the track
    plant0 = np.zeros((3, 7))
    plant0[1, :3] = [0.4, 0.25, 0.05]
    plant0[2, :3] = [-0.3, -0.2, -0.04]
    r = mpc.closed_loop(model, plant0, n_sim, param_id=np.full(3, pid, np.int32))
    assert (r["steps"] == n_sim).all()
    assert (r["exit_hist"] == 0).all()
    for b in range(3):
        # oracle closed loop from the same start state
        orig = cl.run.__defaults__
        h = _oracle_run(fsg.track, model_name, n_sim, plant0[b])
        xs = np.array(h["x"])
        assert xs.shape == (n_sim, 7)
        err = np.abs(r["plant_hist"][b] - xs).max()
        assert err < 2e-6, (b, err)
        assert np.abs(r["n_hist"][b, :n_sim] - np.array(h["n"])[:n_sim]).max() < 2e-6


def _oracle_run(track, model, n_sim, plant0):
    """oracle.closed_loop.run with a custom initial plant state (main.m:59 uses zeros)."""
    from oracle import closed_loop as cl, vehicle as vm, ltv
    import numpy as np
    N_x = 5 if model == "KINEMATIC" else 7
    step = ltv.ltvmpc_kinetmatic_curvilinear if model == "KINEMATIC" else ltv.ltvmpc_dynamic_curvilinear
    x_opt, u_opt = cl.initial_guess(N_x, 2, 40, 0.05)
    x = plant0.copy()
    vs, ss = (0, 0), (0, 0)
    hist = dict(x=[], n=[])
    for i in range(n_sim):
        s, n, mu = vm.cartesian_to_curvilinear(x[0], x[1], x[2], track.x_spline, track.y_spline, track.dl, x_opt[0, 0])
        x0 = np.array([s, n, mu, np.linalg.norm(x[3:5]), x[6]]) if model == "KINEMATIC" else np.array([s, n, mu, x[3], x[4], x[5], x[6]])
        hist["n"].append(n)
        x_ref = cl.make_reference(x0, x[3], N_x, 40, 0.05)
        out = step(x0, x_ref, track.kappa, 0.05, x_opt, u_opt)
        assert out[2] == 0
        x_opt = np.asarray(out[1]).reshape(N_x, 40, order="F")
        u_opt = np.asarray(out[0]).reshape(2, 40, order="F")
        for _ in range(10):
            vr, vs = vm.pid_controller(x_opt[3, 0], x[3], (16000.0, 0, 0, 2800), vs)
            sr, ss = vm.pid_controller(x_opt[N_x - 1, 0], x[6], (80.0, 0, 0, 0.8), ss)
            x = vm.integrate_cart_dyn(x, np.array([vr, sr]), 0.005)
        hist["x"].append(x.copy())
    return hist


def test_obtain_reference_matches_reference_m_file_bit_for_bit(mpc):
    """fsae_obtain_reference_host (util/obtain_reference.m for a batch of vehicles) against the outputs of the
    reference's own .m file on the committed plans, and against the oracle on a larger batch of start points."""
    from conftest import load_golden
    from oracle import reference as rf
    g = load_golden("reference_m_obtain_reference.npz")
    groups = {}
    for i in range(int(g["n"])):
        key = (int(g[f"c{i}_N_s"]), float(g[f"c{i}_ds"]), int(g[f"c{i}_N_t"]), float(g[f"c{i}_dt"]))
        groups.setdefault(key, []).append(i)
    for (N_s, ds, N_t, dt), idx in groups.items():
        x, t = g[f"c{idx[0]}_x"], g[f"c{idx[0]}_t"]
        s0 = np.array([float(g[f"c{i}_s0"]) for i in idx])
        out = mpc.obtain_reference(x.reshape(N_s, 8), ds, t, s0, dt, N_t)
        for j, i in enumerate(idx):
            assert np.array_equal(out[j].T, g[f"c{i}_x_ref"]), (N_s, i)
        rng = np.random.default_rng(N_s)
        s0 = rng.uniform(-0.5 * ds * N_s, 2.5 * ds * N_s, 257)          # ragged batch, before / beyond one lap
        out = mpc.obtain_reference(x.reshape(N_s, 8), ds, t, s0, dt, N_t)
        for j in range(0, 257, 16):
            assert np.array_equal(out[j].T, rf.obtain_reference(x, ds, N_s, t, s0[j], dt, N_t)), j
    assert mpc.obtain_reference(x.reshape(N_s, 8), ds, t, np.zeros(0), dt, N_t).shape == (0, N_t, 7)
