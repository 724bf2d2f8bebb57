"""Parity against the WIDE set of vectors computed by the reference's own .m source
(scripts/make_reference_wide.py -> tests/golden/reference_m_wide.npz): 60 kinematic + 52 dynamic
problems at the default horizon (perturbed start states, hard cases, all three tracks), horizons 20 and
80 for both models, and the euler/rk2/rk4 linearisation schemes of both models.

The two large matrices are stored as probes  xA @ V,  B_bar @ V  (V seeded, in the file).

CPU : the numpy oracle reproduces every reference-computed stage to round-off.
GPU : the CUDA linearise / condense stages and the fused step against the same vectors.
"""
import numpy as np
import pytest

from conftest import load_golden, GoldenTrack, c_layout, DT

TRACKS = ("fsg2019", "fss2019", "fso2020")
GROUPS = [("kinematic", 40), ("kinematic", 20), ("kinematic", 80), ("dynamic", 40), ("dynamic", 20), ("dynamic", 80)]


@pytest.fixture(scope="module")
def w():
    return load_golden("reference_m_wide.npz")


@pytest.fixture(scope="module")
def tracks():
    return [GoldenTrack(n) for n in TRACKS]


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin)
    assert np.array_equal(np.sign(a[~fin]), np.sign(b[~fin]))
    if not fin.any():
        return 0.0
    return float(np.max(np.abs(a[fin] - b[fin])) / max(1.0, np.max(np.abs(b[fin]))))


def grp(w, model, N):
    p = f"{model}_N{N}_"
    return {k[len(p):]: v for k, v in w.items() if k.startswith(p)}


def test_wide_fixture_coverage(w):
    assert grp(w, "kinematic", 40)["x0"].shape[0] >= 50 and grp(w, "dynamic", 40)["x0"].shape[0] >= 50
    for model in ("kinematic", "dynamic"):
        assert len(set(grp(w, model, 40)["track"].tolist())) >= 2
        for N in (20, 80):
            assert grp(w, model, N)["x0"].shape[0] >= 3
    ex = set(w["executed"].tolist())
    for s in ("euler", "rk2", "rk4"):
        assert f"{s}_kinematic_curvilinear" in ex and f"{s}_dynamic_curvilinear" in ex


@pytest.mark.parametrize("model,N", GROUPS)
def test_oracle_reproduces_reference_executed_stages(w, tracks, model, N):
    from oracle import ltv
    g = grp(w, model, N)
    build = ltv.build_kinematic_qp if model == "kinematic" else ltv.build_dynamic_qp
    V = g["V"]
    for b in range(g["x0"].shape[0]):
        tr = tracks[int(g["track"][b])]
        q = build(g["x0"][b], g["x_ref"][b], tr.kappa, DT, g["x_lin"][b], g["u_lin"][b])
        for k in ("A", "B", "d", "A_bar", "d_bar", "H", "f", "lb", "ub", "lbA", "ubA"):
            assert rel(q[k], g[k][b]) < 1e-12, (k, b)
        nU = 2 * N
        assert rel(q["B_bar"][:, :nU] @ V[:nU], g["B_bar_probe"][b]) < 1e-12, b
        assert rel(q["xA"] @ V, g["xA_probe"][b]) < 1e-12, b
        assert rel(np.abs(q["xA"]).sum(axis=1), g["xA_abs_sum"][b]) < 1e-12, b


@pytest.mark.parametrize("model", ["kinematic", "dynamic"])
def test_oracle_linearisation_schemes_match_reference_m_files(w, tracks, model):
    from oracle import ltv
    g = grp(w, model, 40)
    for sch in ("euler", "rk2", "rk4"):
        fn = getattr(ltv, f"{sch}_{model}_curvilinear")
        for j, b in enumerate(w[f"{model}_schemes_idx"]):
            tr = tracks[int(g["track"][b])]
            A, B, d = fn(g["x_lin"][b], g["u_lin"][b], tr.kappa, DT)
            assert rel(A, w[f"{model}_schemes_{sch}_A"][j]) < 1e-12
            assert rel(B, w[f"{model}_schemes_{sch}_B"][j]) < 1e-12
            assert rel(d, w[f"{model}_schemes_{sch}_d"][j]) < 1e-12


@pytest.mark.parametrize("model,N,stride", [("kinematic", 40, 6), ("kinematic", 20, 2), ("dynamic", 40, 8), ("dynamic", 20, 3)])
def test_oracle_step_matches_reference_postprocessing(w, tracks, model, N, stride):
    """ltvmpc_*_curvilinear.m:57-60 around the intercepted qpOASES call (a stride of the set: the
    GPU test covers all of it)."""
    from oracle import ltv
    g = grp(w, model, N)
    step = ltv.ltvmpc_kinetmatic_curvilinear if model == "kinematic" else ltv.ltvmpc_dynamic_curvilinear
    for b in range(0, g["x0"].shape[0], stride):
        tr = tracks[int(g["track"][b])]
        u, x, ef, fv, sl, _ = step(g["x0"][b], g["x_ref"][b], tr.kappa, DT, g["x_lin"][b], g["u_lin"][b])
        assert ef == g["exitflag"][b]
        if ef == 0:
            assert rel(u, g["u_opt"][b]) < 1e-9 and rel(x, g["x_opt"][b]) < 1e-9
            assert abs(fv - g["fval"][b]) < 1e-8 * (1 + abs(fv)) and rel(sl, g["slack"][b]) < 1e-9


# ------------------------------------------------------------------------------- GPU
def _ids(mpc, fm, g, model):
    mid = fm.KINEMATIC if model == "kinematic" else fm.DYNAMIC
    pid = 12 + mid
    mpc.set_params(pid, fm.default_params(mid))
    B = g["x0"].shape[0]
    return mid, dict(track_id=g["track"].astype(np.int32), param_id=np.full(B, pid, np.int32))


@pytest.mark.gpu
@pytest.mark.parametrize("model,N", GROUPS)
def test_cuda_stages_match_wide_reference_vectors(mpc, w, model, N):
    import fsae_mpc_b200 as fm
    g = grp(w, model, N)
    mid, ids = _ids(mpc, fm, g, model)
    V = g["V"]
    nU = 2 * N
    A, Bm, d = mpc.linearise(mid, c_layout(g["x_lin"]), c_layout(g["u_lin"]), DT, **ids)
    assert rel(A, g["A"].transpose(0, 3, 1, 2)) < 1e-11
    assert rel(Bm, g["B"].transpose(0, 3, 1, 2)) < 1e-11
    assert rel(d, g["d"].transpose(0, 2, 1)) < 1e-11
    o = mpc.condense(mid, g["x0"], c_layout(g["x_ref"]), DT, c_layout(g["x_lin"]), c_layout(g["u_lin"]), **ids)
    for k in ("A_bar", "d_bar", "H", "f", "lb", "ub", "lbA", "ubA"):
        assert rel(o[k], g[k]) < 1e-10, k
    assert rel(np.einsum("bij,jk->bik", o["B_bar"][:, :, :nU], V[:nU]), g["B_bar_probe"]) < 1e-10
    assert rel(np.einsum("bij,jk->bik", o["xA"], V), g["xA_probe"]) < 1e-10
    assert rel(np.abs(o["xA"]).sum(axis=2), g["xA_abs_sum"]) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("model,N", GROUPS)
def test_cuda_fused_step_matches_wide_reference_vectors(mpc, w, model, N):
    """|du|inf <= 1e-6 relative (north_star) against the reference's post-processed outputs."""
    import fsae_mpc_b200 as fm
    g = grp(w, model, N)
    mid, ids = _ids(mpc, fm, g, model)
    step = mpc.ltvmpc_kinetmatic_curvilinear if model == "kinematic" else mpc.ltvmpc_dynamic_curvilinear
    r = step(g["x0"], c_layout(g["x_ref"]), DT, c_layout(g["x_lin"]), c_layout(g["u_lin"]), **ids)
    assert np.array_equal(r.exitflag, g["exitflag"].astype(np.int32))
    ok = g["exitflag"] == 0
    assert ok.sum() >= 0.9 * ok.size
    assert rel(r.u_opt[ok], g["u_opt"][ok]) < 1e-6 and rel(r.x_opt[ok], g["x_opt"][ok]) < 1e-6
    assert np.max(np.abs(r.fval[ok] - g["fval"][ok]) / (1 + np.abs(g["fval"][ok]))) < 1e-7
    assert rel(r.slack_opt[ok], g["slack"][ok]) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["kinematic", "dynamic"])
def test_cuda_linearisation_schemes_match_reference_m_files(mpc, w, model):
    import fsae_mpc_b200 as fm
    g = grp(w, model, 40)
    mid = fm.KINEMATIC if model == "kinematic" else fm.DYNAMIC
    idx = w[f"{model}_schemes_idx"]
    for pid, (sch, code) in enumerate((("euler", 1), ("rk2", 2), ("rk4", 4)), start=20):
        p = fm.default_params(mid)
        p.lin_scheme = code
        mpc.set_params(pid, p)
        A, Bm, d = mpc.linearise(mid, c_layout(g["x_lin"][idx]), c_layout(g["u_lin"][idx]), DT,
                                 track_id=g["track"][idx].astype(np.int32), param_id=np.full(len(idx), pid, np.int32))
        assert rel(A, w[f"{model}_schemes_{sch}_A"].transpose(0, 3, 1, 2)) < 1e-11, sch
        assert rel(Bm, w[f"{model}_schemes_{sch}_B"].transpose(0, 3, 1, 2)) < 1e-11, sch
        assert rel(d, w[f"{model}_schemes_{sch}_d"].transpose(0, 2, 1)) < 1e-11, sch
