"""GPU: the dense qpOASES drop-in (fsae_qpoases_host) against the oracle QP solver -- on the
condensed MPC QPs exactly as the reference hands them to qpOASES (reference-executed H, f, xA,
lb, ub, lbA, ubA) and on random QPs with bounds, two-sided rows and infeasible cases."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def test_dense_qp_on_reference_condensed_kinematic(mpc):
    from oracle import qp
    g = load_golden("reference_m_kinematic_fsg2019.npz")
    o = mpc.qpOASES(g["H"], g["f"], g["xA"], g["lb"], g["ub"], g["lbA"], g["ubA"])
    assert (o["exitflag"] == 0).all()
    nU = g["u_opt"].shape[1]
    scale = np.maximum(1.0, np.abs(g["u_opt"]).max(axis=1))
    assert (np.abs(o["x"][:, :nU] - g["u_opt"]).max(axis=1) / scale).max() < 1e-6
    assert np.abs(o["x"][:, nU:] - g["slack"]).max() < 1e-7
    for b in range(g["H"].shape[0]):
        sol = qp.qpoases(g["H"][b], g["f"][b], g["xA"][b], g["lb"][b], g["ub"][b], g["lbA"][b], g["ubA"][b])
        assert abs(o["fval"][b] - sol.fval) < 1e-7 * (1 + abs(sol.fval))
        assert np.array_equal(o["workingSetB"][b], sol.workingSetB)
        assert np.array_equal(o["workingSetC"][b], sol.workingSetC)
        k = qp.kkt_residuals(g["H"][b], g["f"][b], g["xA"][b], g["lb"][b], g["ub"][b], g["lbA"][b], g["ubA"][b],
                             o["x"][b], o["lam"][b])
        assert max(k["primal"], k["stationarity"], k["complementarity"]) < 1e-7


def test_dense_qp_on_reference_condensed_dynamic(mpc):
    g = load_golden("reference_m_dynamic_fss2019.npz")
    o = mpc.qpOASES(g["H"], g["f"], g["xA"], g["lb"], g["ub"], g["lbA"], g["ubA"])
    assert (o["exitflag"] == 0).all()
    nU = g["u_opt"].shape[1]
    scale = np.maximum(1.0, np.abs(g["u_opt"]).max(axis=1))
    assert (np.abs(o["x"][:, :nU] - g["u_opt"]).max(axis=1) / scale).max() < 1e-6


def test_dense_qp_random_and_infeasible(mpc):
    from oracle import qp
    rng = np.random.default_rng(11)
    B, n, m = 24, 12, 20
    H = np.empty((B, n, n)); g = rng.normal(size=(B, n)) * 3
    A = rng.normal(size=(B, m, n))
    for b in range(B):
        G = rng.normal(size=(n, n))
        H[b] = G @ G.T + 0.5 * np.eye(n)
    lb, ub = -np.ones((B, n)), np.ones((B, n))
    lbA, ubA = -0.5 * np.ones((B, m)), 0.8 * np.ones((B, m))
    # last two problems: contradictory rows -> infeasible
    A[-1, 0] = 0; A[-1, 0, 0] = 1; lbA[-1, 0] = 5; ubA[-1, 0] = 6
    A[-2, 1] = A[-2, 0]; lbA[-2, 0] = 0.3; ubA[-2, 0] = 0.4; lbA[-2, 1] = -0.4; ubA[-2, 1] = -0.3
    o = mpc.qpOASES(H, g, A, lb, ub, lbA, ubA)
    for b in range(B):
        sol = qp.qpoases(H[b], g[b], A[b], lb[b], ub[b], lbA[b], ubA[b])
        assert o["exitflag"][b] == sol.exitflag, b
        if sol.exitflag == 0:
            assert np.abs(o["x"][b] - sol.x).max() < 1e-8
            assert np.abs(o["lam"][b] - sol.lam).max() < 1e-6 * (1 + np.abs(sol.lam).max())
    assert o["exitflag"][-1] == -2 and o["exitflag"][-2] == -2


def test_dense_qp_bounds_only_and_too_large(mpc):
    import fsae_mpc_b200 as fm
    rng = np.random.default_rng(3)
    B, n = 4, 7
    H = np.tile(np.eye(n) * 2, (B, 1, 1)); g = rng.normal(size=(B, n)) * 4
    o = mpc.qpOASES(H, g, None, -np.ones((B, n)), np.ones((B, n)), None, None)
    assert np.allclose(o["x"], np.clip(-g / 2, -1, 1), atol=1e-12)
    with pytest.raises(fm.FsaeError):
        mpc.qpOASES(np.tile(np.eye(200), (1, 1, 1)), np.zeros((1, 200)), None, -np.ones((1, 200)), np.ones((1, 200)), None, None)
    # 96 <= nV <= 191: the large instantiation (operator tile partly in shared memory, H in a global slab)
    n = 120
    H = np.tile(np.eye(n) * 2, (B, 1, 1)); g = rng.normal(size=(B, n)) * 4
    o = mpc.qpOASES(H, g, None, -np.ones((B, n)), np.ones((B, n)), None, None)
    assert (o["exitflag"] == 0).all() and np.allclose(o["x"], np.clip(-g / 2, -1, 1), atol=1e-12)


@pytest.mark.parametrize("model", ["kinematic", "dynamic"])
def test_dense_qp_on_horizon_80_condensed_qps(mpc, model):
    """The literal qpOASES(H,f,xA,lb,ub,lbA,ubA) drop-in on the condensed QPs of horizon 80 (nV = 161 / 164, 480 / 1600
    rows): the large instantiation of the dense kernel must return the minimiser the fused step returns for the same
    problems, and a KKT point of the QP it was handed."""
    import fsae_mpc_b200 as fm
    from conftest import c_layout, DT
    from oracle import qp
    mid = fm.KINEMATIC if model == "kinematic" else fm.DYNAMIC
    g = load_golden("kinematic_lap_fsg2019_N80.npz" if model == "kinematic" else "dynamic_lap_fss2019_N80.npz")
    pick = [1, 6, 11]
    tid = np.full(len(pick), 0 if model == "kinematic" else 1, np.int32)
    pid = np.full(len(pick), 50 + mid, np.int32)
    mpc.set_params(50 + mid, fm.default_params(mid))
    args = (g["x0"][pick], c_layout(g["x_ref"][pick]), DT, c_layout(g["x_lin"][pick]), c_layout(g["u_lin"][pick]))
    q = mpc.condense(mid, *args, track_id=tid, param_id=pid)
    step = mpc.ltvmpc_kinetmatic_curvilinear if model == "kinematic" else mpc.ltvmpc_dynamic_curvilinear
    r = step(*args, track_id=tid, param_id=pid)
    o = mpc.qpOASES(q["H"], q["f"], q["xA"], q["lb"], q["ub"], q["lbA"], q["ubA"])
    assert (o["exitflag"] == 0).all() and (r.exitflag == 0).all()
    nU = r.u_opt.shape[1]
    scale = np.maximum(1.0, np.abs(r.u_opt).max(axis=1))
    assert (np.abs(o["x"][:, :nU] - r.u_opt).max(axis=1) / scale).max() < 1e-6
    assert np.abs(o["x"][:, nU:] - r.slack_opt).max() < 1e-6
    assert np.max(np.abs(o["fval"] + q["const"] - r.fval) / (1 + np.abs(r.fval))) < 1e-7
    for b in range(len(pick)):
        k = qp.kkt_residuals(q["H"][b], q["f"][b], q["xA"][b], q["lb"][b], q["ub"][b], q["lbA"][b], q["ubA"][b], o["x"][b], o["lam"][b])
        assert max(k["primal"], k["stationarity"], k["complementarity"]) < 1e-7, (b, k)


def test_dense_qp_flat_variable_on_its_upper_bound(mpc):
    """A zero-curvature variable with NEGATIVE linear cost starts on its UPPER bound (normal -e_i): x, the
    multipliers and the working set must come out right, with and without the coupling row active."""
    from oracle import qp
    rng = np.random.default_rng(5)
    B, n, m = 16, 6, 3
    H = np.zeros((B, n, n)); g = rng.normal(size=(B, n)) * 2
    A = rng.normal(size=(B, m, n))
    for b in range(B):
        G = rng.normal(size=(n - 2, n - 2))
        H[b, :n - 2, :n - 2] = G @ G.T + 0.5 * np.eye(n - 2)
    g[:, n - 1] = -np.abs(g[:, n - 1]) - 0.5          # flat, pushed up: starts on ub
    g[:, n - 2] = np.abs(g[:, n - 2]) + 0.5           # flat, pushed down: starts on lb
    lb, ub = -np.ones((B, n)), np.ones((B, n))
    ub[:, n - 1] = rng.uniform(0.5, 2.0, B)
    A[:, 0, n - 1] = 1.0                              # coupling row that involves the flat variable
    lbA, ubA = -2.0 * np.ones((B, m)), 2.0 * np.ones((B, m))
    ubA[::2, 0] = 0.2                                 # every other problem: the row forces it off its bound
    o = mpc.qpOASES(H, g, A, lb, ub, lbA, ubA)
    n_off = 0
    for b in range(B):
        sol = qp.qpoases(H[b], g[b], A[b], lb[b], ub[b], lbA[b], ubA[b])
        assert sol.exitflag == 0 and o["exitflag"][b] == 0, b
        k = qp.kkt_residuals(H[b], g[b], A[b], lb[b], ub[b], lbA[b], ubA[b], o["x"][b], o["lam"][b])
        assert max(k["primal"], k["stationarity"], k["complementarity"]) < 1e-7, (b, k)
        assert np.abs(o["x"][b] - sol.x).max() < 1e-7, b
        assert np.abs(o["lam"][b] - sol.lam).max() < 1e-6 * (1 + np.abs(sol.lam).max()), b
        assert np.array_equal(o["workingSetB"][b], sol.workingSetB) and np.array_equal(o["workingSetC"][b], sol.workingSetC), b
        n_off += o["x"][b, n - 1] < ub[b, n - 1] - 1e-9
    assert (o["workingSetB"][:, n - 1] == 1).any() and n_off >= 2
