"""GPU parity tests of the curvilinear DYNAMIC model path (BASELINE configs[2]: tyre forces,
7 states, 4 slack variables, 800 general constraints) against the oracle / golden fixtures."""
import numpy as np
import pytest

from conftest import load_golden, c_layout, DT

pytestmark = pytest.mark.gpu
U_RTOL = 1e-6


def rel(a, b):
    return np.max(np.abs(a - b)) / max(1.0, np.max(np.abs(b)))


@pytest.mark.parametrize("scheme", [1, 2, 4])
def test_linearise_matches_oracle(mpc, fss, scheme):
    import fsae_mpc_b200 as fm
    from oracle import ltv
    g = load_golden("dynamic_lap_fss2019.npz")
    p = fm.default_params(fm.DYNAMIC)
    p.lin_scheme = scheme
    mpc.set_params(2, p)
    B = 6
    A, Bm, d = mpc.linearise(fm.DYNAMIC, c_layout(g["x_lin"][:B]), c_layout(g["u_lin"][:B]), DT,
                             track_id=np.ones(B, np.int32), param_id=np.full(B, 2, np.int32))
    fn = {1: ltv.euler_dynamic_curvilinear, 2: ltv.rk2_dynamic_curvilinear, 4: ltv.rk4_dynamic_curvilinear}[scheme]
    for b in range(B):
        Ao, Bo, do = fn(g["x_lin"][b], g["u_lin"][b], fss.kappa, DT)
        assert rel(A[b], Ao.transpose(2, 0, 1)) < 1e-10
        assert rel(Bm[b], Bo.transpose(2, 0, 1)) < 1e-10
        assert rel(d[b], do.T) < 1e-10


def _dyn_params(mpc):
    import fsae_mpc_b200 as fm
    mpc.set_params(3, fm.default_params(fm.DYNAMIC))
    return 3


def test_condense_matches_golden_stage(mpc):
    import fsae_mpc_b200 as fm
    g = load_golden("dynamic_lap_fss2019.npz")
    idx = g["stage_idx"]
    B = len(idx)
    pid = _dyn_params(mpc)
    o = mpc.condense(fm.DYNAMIC, g["x0"][idx], c_layout(g["x_ref"][idx]), DT, c_layout(g["x_lin"][idx]),
                     c_layout(g["u_lin"][idx]), track_id=np.ones(B, np.int32), param_id=np.full(B, pid, np.int32))
    for k in ("A_bar", "B_bar", "d_bar", "H", "f", "xA", "const"):
        assert rel(o[k], g["stage_" + k]) < 1e-9, k
    for k in ("lbA", "ubA", "lb", "ub"):
        a, b = o[k], g["stage_" + k]
        assert np.array_equal(np.isinf(a), np.isinf(b)), k
        fin = np.isfinite(b)
        assert np.array_equal(np.sign(a[~fin]), np.sign(b[~fin])), k
        assert np.max(np.abs(a[fin] - b[fin]) / (1 + np.abs(b[fin]))) < 1e-9, k


def _check(r, g):
    assert np.array_equal(r.exitflag, g["exitflag"].astype(np.int32))
    scale = np.maximum(1.0, np.max(np.abs(g["u_opt"]), axis=1))
    du = np.max(np.abs(r.u_opt - g["u_opt"]), axis=1) / scale
    assert du.max() <= U_RTOL, f"max |du|inf rel = {du.max():.3e} at {du.argmax()}"
    xs = np.maximum(1.0, np.max(np.abs(g["x_opt"]), axis=1))
    assert (np.max(np.abs(r.x_opt - g["x_opt"]), axis=1) / xs).max() <= U_RTOL
    assert np.max(np.abs(r.fval - g["fval"]) / (1 + np.abs(g["fval"]))) <= 1e-7
    assert np.max(np.abs(r.slack_opt - g["slack"])) <= 1e-7
    same = (r.workingSetB == g["wsB"]).all(axis=1) & (r.workingSetC == g["wsC"]).all(axis=1)
    assert same.mean() >= 0.95, f"working set differs on {(~same).sum()} of {same.size} problems"


@pytest.mark.parametrize("fixture,tid", [("dynamic_lap_fss2019.npz", 1), ("dynamic_perturbed_fss2019.npz", 1),
                                         ("dynamic_lap_fsg2019.npz", 0)])
def test_fused_step_matches_golden(mpc, fixture, tid):
    g = load_golden(fixture)
    B = g["x0"].shape[0]
    pid = _dyn_params(mpc)
    r = mpc.ltvmpc_dynamic_curvilinear(g["x0"], c_layout(g["x_ref"]), DT, c_layout(g["x_lin"]), c_layout(g["u_lin"]),
                                       track_id=np.full(B, tid, np.int32), param_id=np.full(B, pid, np.int32))
    _check(r, g)


def test_hard_cases_match_oracle(mpc):
    """The problems of the synthetic bench batches with the longest pivot sequences (hundreds of partial steps) and
    the two on which earlier kernel variants stopped at a point with a negative recomputed multiplier
    (|du| ~ 1e-2): scripts/make_hard_cases.py.  Kinematic and dynamic model, oracle solutions."""
    import fsae_mpc_b200 as fm
    g = load_golden("hard_cases.npz")
    pid = _dyn_params(mpc)
    for tag, model, tid in (("dyn", fm.DYNAMIC, 1), ("kin", fm.KINEMATIC, 0)):
        B = g[tag + "_x0"].shape[0]
        step = mpc.ltvmpc_dynamic_curvilinear if model == fm.DYNAMIC else mpc.ltvmpc_kinetmatic_curvilinear
        kw = dict(track_id=np.full(B, tid, np.int32))
        if model == fm.DYNAMIC:
            kw["param_id"] = np.full(B, pid, np.int32)
        r = step(g[tag + "_x0"], g[tag + "_x_ref"], DT, g[tag + "_x_lin"], g[tag + "_u_lin"], **kw)
        assert np.array_equal(r.exitflag, g[tag + "_exitflag"].astype(np.int32)), tag
        scale = np.maximum(1.0, np.max(np.abs(g[tag + "_u_opt"]), axis=1))
        du = np.max(np.abs(r.u_opt - g[tag + "_u_opt"]), axis=1) / scale
        assert du.max() <= U_RTOL, f"{tag}: max |du|inf rel = {du.max():.3e} at {du.argmax()}"
        assert np.max(np.abs(r.fval - g[tag + "_fval"]) / (1 + np.abs(g[tag + "_fval"]))) <= 1e-9, tag
        assert np.max(np.abs(r.slack_opt - g[tag + "_slack"])) <= 1e-7, tag
