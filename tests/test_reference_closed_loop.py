"""Closed loop and plant side against the REFERENCE'S OWN SOURCE: main.m's simulation loop
(main.m:91-190) and cartesian_to_curvilinear.m, closest_point.m, pid_controller.m,
integrate_cart_dyn.m, f_cart_dyn.m executed by oracle/mlab
(scripts/make_reference_closed_loop.py -> tests/golden/reference_m_closed_loop.npz).

CPU : oracle/vehicle.py (plant functions) and oracle/closed_loop.py (main.m restatement) reproduce them.
GPU : fsae_closed_loop_host reproduces main.m's histories (marked gpu).
"""
import numpy as np
import pytest

from conftest import load_golden, GoldenTrack

TRACKS = ("fsg2019", "fss2019", "fso2020")


@pytest.fixture(scope="module")
def g():
    return load_golden("reference_m_closed_loop.npz")


def test_fixture_was_produced_by_the_reference_files(g):
    ex = set(g["executed"].tolist())
    for f in ("cartesian_to_curvilinear", "closest_point", "pid_controller", "integrate_cart_dyn", "f_cart_dyn",
              "ltvmpc_kinetmatic_curvilinear", "ltvmpc_dynamic_curvilinear", "interpolate_angle"):
        assert f in ex, f


def test_plant_model_matches_reference_m_files(g):
    from oracle import vehicle as vm
    for i, (x, u) in enumerate(zip(g["plant_x"], g["plant_u"])):
        f = vm.f_cart_dyn(x, u)
        assert np.max(np.abs(f - g["f_cart_dyn"][i])) <= 1e-12 * max(1.0, np.max(np.abs(f))), i
        for j, h in enumerate(g["integrate_dt"]):
            xn = vm.integrate_cart_dyn(x, u, float(h))
            assert np.max(np.abs(xn - g["integrate_cart_dyn"][j, i])) <= 1e-12 * max(1.0, np.max(np.abs(xn))), (i, j)


def test_pid_controller_matches_reference_m_file(g):
    from oracle import vehicle as vm
    status = (0.0, 0.0)
    for i, (row, out) in enumerate(zip(g["pid_in"], g["pid_out"])):
        if i % 8 == 0:
            status = (0.0, 0.0)                   # the fixture chains 8 calls per settings row
        o, status = vm.pid_controller(row[4], row[5], tuple(row[:4]), status)
        assert abs(o - out[0]) <= 1e-12 * max(1.0, abs(out[0]))
        assert abs(status[0] - out[1]) <= 1e-12 * max(1.0, abs(out[1])) and abs(status[1] - out[2]) <= 1e-12
    assert (np.abs(g["pid_out"][:, 0]) == g["pid_in"][:, 3]).any()        # saturation is covered


def test_projection_onto_track_matches_reference_m_files(g):
    from oracle import vehicle as vm
    tr = [GoldenTrack(n) for n in TRACKS]
    wrapped = 0
    for row, out in zip(g["c2c_in"], g["c2c_out"]):
        t = tr[int(row[0])]
        s, n, mu = vm.cartesian_to_curvilinear(row[1], row[2], row[3], t.x_spline, t.y_spline, t.dl, row[4])
        assert abs(s - out[0]) <= 1e-10 and abs(n - out[1]) <= 1e-10 and abs(mu - out[2]) <= 1e-10
        wrapped += abs(row[3]) > np.pi
    assert wrapped >= 10                                                   # angdiff's wrap is covered


@pytest.mark.parametrize("model,n_cmp", [("KINEMATIC", 24), ("DYNAMIC", 12)])
def test_oracle_closed_loop_matches_main_m(g, model, n_cmp):
    """oracle/closed_loop.run == main.m:91-190 (first n_cmp steps of the recorded run; the GPU test
    covers the whole record)."""
    from oracle import closed_loop as cl
    tr = GoldenTrack(str(g[f"main_{model}_track"]))
    h = cl.run(tr.track, model, n_sim=n_cmp)
    assert h["steps"] == n_cmp and all(e == 0 for e in h["exitflag"])
    assert (g[f"main_{model}_exit_status"][:n_cmp] == 0).all()
    xs = np.array(h["x"])
    assert np.max(np.abs(xs - g[f"main_{model}_x_history"][:n_cmp])) < 1e-7
    assert np.max(np.abs(np.array(h["n"])[:n_cmp] - g[f"main_{model}_n_list"][:n_cmp])) < 1e-8
    assert np.max(np.abs(np.array(h["u0"]) - g[f"main_{model}_u_opt_history"][:n_cmp])) < 1e-6
    fv = np.array(h["fval"])
    assert np.max(np.abs(fv - g[f"main_{model}_objective"][:n_cmp]) / (1 + np.abs(fv))) < 1e-7


@pytest.mark.gpu
@pytest.mark.parametrize("model_name,tid", [("KINEMATIC", 0), ("DYNAMIC", 1)])
def test_cuda_closed_loop_matches_main_m(mpc, g, model_name, tid):
    """fsae_closed_loop_host from main.m's own start state against main.m's recorded histories."""
    import fsae_mpc_b200 as fm
    model = fm.KINEMATIC if model_name == "KINEMATIC" else fm.DYNAMIC
    assert TRACKS[tid] == str(g[f"main_{model_name}_track"])
    n_sim = int(g[f"main_{model_name}_steps"])
    pid = 10 + model
    mpc.set_params(pid, fm.default_params(model))
    r = mpc.closed_loop(model, np.zeros((1, 7)), n_sim, track_id=np.full(1, tid, np.int32), param_id=np.full(1, pid, np.int32))
    assert r["steps"][0] == n_sim and (r["exit_hist"] == 0).all()
    assert np.max(np.abs(r["plant_hist"][0] - g[f"main_{model_name}_x_history"])) < 2e-6
    assert np.max(np.abs(r["n_hist"][0, :n_sim] - g[f"main_{model_name}_n_list"][:n_sim])) < 2e-6
    assert np.max(np.abs(r["plant"][0] - g[f"main_{model_name}_x_final"])) < 2e-6
