"""Parity against vectors computed by the REFERENCE'S OWN .m SOURCE (executed by oracle/mlab,
scripts/make_reference_fixtures.py -> tests/golden/reference_m_*.npz).

CPU part : the numpy oracle reproduces every reference-computed stage (linearise,
           sequential_integration, constraints, generate_qp) to round-off, and -- where
           /root/reference is mounted -- the .m files are re-executed live.
GPU part : the CUDA stages and the fused step against the same vectors (marked gpu).
"""
import os

import numpy as np
import pytest

from conftest import load_golden, GoldenTrack, c_layout, DT

CASES = [("kinematic", "fsg2019"), ("dynamic", "fss2019")]


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin)
    assert np.array_equal(np.sign(a[~fin]), np.sign(b[~fin]))
    if not fin.any():
        return 0.0
    return float(np.max(np.abs(a[fin] - b[fin])) / max(1.0, np.max(np.abs(b[fin]))))


@pytest.mark.parametrize("model,track", CASES)
def test_oracle_reproduces_reference_executed_stages(model, track):
    from oracle import ltv
    g = load_golden(f"reference_m_{model}_{track}.npz")
    tr = GoldenTrack(track)
    build = ltv.build_kinematic_qp if model == "kinematic" else ltv.build_dynamic_qp
    for b in range(g["x0"].shape[0]):
        q = build(g["x0"][b], g["x_ref"][b], tr.kappa, DT, g["x_lin"][b], g["u_lin"][b])
        for k in ("A", "B", "d", "A_bar", "d_bar", "H", "f", "xA", "lb", "ub", "lbA", "ubA"):
            assert rel(q[k], g[k][b]) < 1e-12, (k, b)
        # the reference's B_bar is the un-augmented one (sequential_integration.m output)
        assert rel(q["B_bar"][:, :g["B_bar"].shape[2]], g["B_bar"][b]) < 1e-12


@pytest.mark.parametrize("model,track", CASES)
def test_oracle_step_matches_reference_postprocessing(model, track):
    """u_opt / x_opt / fval / slack as computed by ltvmpc_*_curvilinear.m:57-60 around the
    intercepted qpOASES call."""
    from oracle import ltv
    g = load_golden(f"reference_m_{model}_{track}.npz")
    tr = GoldenTrack(track)
    step = ltv.ltvmpc_kinetmatic_curvilinear if model == "kinematic" else ltv.ltvmpc_dynamic_curvilinear
    for b in range(g["x0"].shape[0]):
        u, x, ef, fv, sl, _ = step(g["x0"][b], g["x_ref"][b], tr.kappa, DT, g["x_lin"][b], g["u_lin"][b])
        assert ef == 0 == g["exitflag"][b]
        assert rel(u, g["u_opt"][b]) < 1e-9 and rel(x, g["x_opt"][b]) < 1e-9
        assert abs(fv - g["fval"][b]) < 1e-8 * (1 + abs(fv)) and rel(sl, g["slack"][b]) < 1e-9


@pytest.mark.skipif(not os.path.isdir("/root/reference/mpc/ltv"), reason="reference tree not mounted (GPU box)")
def test_live_execution_of_reference_sources():
    """Re-run the reference .m files now (build container only) -- guards the fixtures and the
    interpreter against drift."""
    from oracle.mlab.interp import Matlab
    from oracle import spline as sp, vehicle as vm
    R = "/root/reference"
    ml = Matlab([R + "/spline", R + "/vehicle_models/curvilinear_kinematic", R + "/vehicle_models/curvilinear_dynamic",
                 R + "/mpc/ltv", R + "/mpc/ltv/kinematic", R + "/mpc/ltv/dynamic"])
    tr = GoldenTrack("fsg2019")
    s = np.array([0.3, 17.2, 250.0, 400.0, -3.0, tr.L - 1e-6])
    k_ml = ml.call("interpolate_curvature", s, tr.x_spline, tr.y_spline, tr.dl)
    assert np.max(np.abs(k_ml.ravel() - tr.kappa(s))) < 1e-14
    kappa = lambda s_, nargout=1: [ml.call("interpolate_curvature", s_, tr.x_spline, tr.y_spline, tr.dl)]
    x = np.array([12.0, 0.2, 0.05, 9.0, 0.1])
    u = np.array([1.0, 0.2])
    assert np.max(np.abs(ml.call("A_curv_kin", x.reshape(-1, 1), u.reshape(-1, 1), kappa)
                         - vm.A_curv_kin(x, u, tr.kappa))) < 1e-14
    xd = np.array([12.0, 0.2, 0.05, 9.0, 0.3, 0.1, 0.05])
    A_ml = ml.call("A_curv_dyn", xd.reshape(-1, 1), u.reshape(-1, 1), kappa, nargout=9)
    A_or = vm.A_curv_dyn(xd, u, tr.kappa)
    assert np.max(np.abs(A_ml[0] - A_or[0])) < 1e-12
    for a, b in zip(A_ml[1:], A_or[1:]):
        assert abs(float(np.asarray(a).ravel()[0]) - b) < 1e-10 * (1 + abs(b))
    g = load_golden("reference_m_kinematic_fsg2019.npz")
    A, B, d = ml.call("rk4_kinematic_curvilinear", g["x_lin"][0], g["u_lin"][0], kappa, DT, nargout=3)
    from oracle import ltv
    Ao, Bo, do = ltv.rk4_kinematic_curvilinear(g["x_lin"][0], g["u_lin"][0], tr.kappa, DT)
    assert max(np.max(np.abs(A - Ao)), np.max(np.abs(B - Bo)), np.max(np.abs(d - do))) < 1e-13
    # spline construction (main.m:15-17) on a coarse resample: periodic spline + reparametrisation
    P = np.array([0.0, 3.0, 5.0, 4.0, 1.0, -1.0]).reshape(-1, 1)
    assert np.max(np.abs(ml.call("make_spline_periodic", P) - sp.make_spline_periodic(P))) < 1e-14


# ------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("model,track,tid", [("kinematic", "fsg2019", 0), ("dynamic", "fss2019", 1)])
def test_cuda_stages_match_reference_executed_vectors(mpc, model, track, tid):
    import fsae_mpc_b200 as fm
    g = load_golden(f"reference_m_{model}_{track}.npz")
    mid = fm.KINEMATIC if model == "kinematic" else fm.DYNAMIC
    B = g["x0"].shape[0]
    pid = 4 + mid
    mpc.set_params(pid, fm.default_params(mid))
    ids = dict(track_id=np.full(B, tid, np.int32), param_id=np.full(B, pid, np.int32))
    A, Bm, d = mpc.linearise(mid, c_layout(g["x_lin"]), c_layout(g["u_lin"]), DT, **ids)
    assert rel(A, g["A"].transpose(0, 3, 1, 2)) < 1e-11
    assert rel(Bm, g["B"].transpose(0, 3, 1, 2)) < 1e-11
    assert rel(d, g["d"].transpose(0, 2, 1)) < 1e-11
    o = mpc.condense(mid, g["x0"], c_layout(g["x_ref"]), DT, c_layout(g["x_lin"]), c_layout(g["u_lin"]), **ids)
    for k in ("A_bar", "d_bar", "H", "f", "xA", "lb", "ub", "lbA", "ubA"):
        assert rel(o[k], g[k]) < 1e-10, k
    assert rel(o["B_bar"][:, :, :g["B_bar"].shape[2]], g["B_bar"]) < 1e-10
    step = mpc.ltvmpc_kinetmatic_curvilinear if model == "kinematic" else mpc.ltvmpc_dynamic_curvilinear
    r = step(g["x0"], c_layout(g["x_ref"]), DT, c_layout(g["x_lin"]), c_layout(g["u_lin"]), **ids)
    assert (r.exitflag == 0).all()
    assert rel(r.u_opt, g["u_opt"]) < 1e-6 and rel(r.x_opt, g["x_opt"]) < 1e-6
    assert np.max(np.abs(r.fval - g["fval"]) / (1 + np.abs(g["fval"]))) < 1e-7
