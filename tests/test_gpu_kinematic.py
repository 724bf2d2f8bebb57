"""GPU parity tests (through the C-ABI) of the kinematic LTV-MPC path against the oracle
and the committed golden fixtures.  Tolerance: north_star's |du|inf <= 1e-6 relative, and
the same active set; the intermediate (pre-QP) stages are held to 1e-9 relative."""
import numpy as np
import pytest

from conftest import load_golden, c_layout, DT

pytestmark = pytest.mark.gpu

U_RTOL = 1e-6


def rel(a, b):
    return np.max(np.abs(a - b)) / max(1.0, np.max(np.abs(b)))


def test_curvature_matches_oracle(mpc, fsg):
    s = np.concatenate([np.linspace(-50, 700, 4001), [0.0, fsg.L, fsg.dl, 2 * fsg.L - 1e-9]])
    k_gpu = mpc.interpolate_curvature(s, 0)
    k_ref = fsg.kappa(s)
    assert np.max(np.abs(k_gpu - k_ref)) <= 1e-12 * max(1.0, np.max(np.abs(k_ref)))


@pytest.mark.parametrize("scheme", [1, 2, 4])
def test_linearise_matches_oracle(mpc, fsg, scheme):
    import fsae_mpc_b200 as fm
    from oracle import ltv
    g = load_golden("kinematic_lap_fsg2019.npz")
    p = fm.default_params(fm.KINEMATIC)
    p.lin_scheme = scheme
    mpc.set_params(1, p)
    B = min(8, g["x_lin"].shape[0])
    pid = np.ones(B, np.int32)
    A, Bm, d = mpc.linearise(fm.KINEMATIC, c_layout(g["x_lin"][:B]), c_layout(g["u_lin"][:B]), DT, param_id=pid)
    fn = {1: ltv.euler_kinematic_curvilinear, 2: ltv.rk2_kinematic_curvilinear, 4: ltv.rk4_kinematic_curvilinear}[scheme]
    for b in range(B):
        Ao, Bo, do = fn(g["x_lin"][b], g["u_lin"][b], fsg.kappa, DT)
        assert rel(A[b], Ao.transpose(2, 0, 1)) < 1e-11
        assert rel(Bm[b], Bo.transpose(2, 0, 1)) < 1e-11
        assert rel(d[b], do.T) < 1e-11


def test_condense_matches_golden_stage(mpc):
    import fsae_mpc_b200 as fm
    g = load_golden("kinematic_lap_fsg2019.npz")
    idx = g["stage_idx"]
    o = mpc.condense(fm.KINEMATIC, g["x0"][idx], c_layout(g["x_ref"][idx]), DT,
                     c_layout(g["x_lin"][idx]), c_layout(g["u_lin"][idx]))
    for k in ("A_bar", "B_bar", "d_bar", "H", "f", "xA", "const"):
        assert rel(o[k], g["stage_" + k]) < 1e-10, k
    for k in ("lbA", "ubA", "lb", "ub"):
        a, b = o[k], g["stage_" + k]
        assert np.array_equal(np.isinf(a), np.isinf(b)), k
        fin = np.isfinite(b)
        assert np.array_equal(np.sign(a[~fin]), np.sign(b[~fin])), k
        assert np.max(np.abs(a[fin] - b[fin]) / (1 + np.abs(b[fin]))) < 1e-10, k


def _check_solution(r, g, sl=slice(None)):
    assert np.array_equal(r.exitflag, g["exitflag"][sl].astype(np.int32))
    scale = np.maximum(1.0, np.max(np.abs(g["u_opt"][sl]), axis=1))
    du = np.max(np.abs(r.u_opt - g["u_opt"][sl]), axis=1) / scale
    assert du.max() <= U_RTOL, f"max |du|inf rel = {du.max():.3e} at {du.argmax()}"
    xs = np.maximum(1.0, np.max(np.abs(g["x_opt"][sl]), axis=1))
    assert (np.max(np.abs(r.x_opt - g["x_opt"][sl]), axis=1) / xs).max() <= U_RTOL
    assert np.max(np.abs(r.fval - g["fval"][sl]) / (1 + np.abs(g["fval"][sl]))) <= 1e-7
    assert np.max(np.abs(r.slack_opt - g["slack"][sl])) <= 1e-7
    # same active set (non-degenerate problems): compare working sets exactly
    same_B = (r.workingSetB == g["wsB"][sl]).all(axis=1)
    same_C = (r.workingSetC == g["wsC"][sl]).all(axis=1)
    assert (same_B & same_C).mean() >= 0.98, f"working set differs on {(~(same_B & same_C)).sum()} problems"


def test_fused_step_matches_golden_lap(mpc):
    g = load_golden("kinematic_lap_fsg2019.npz")
    r = mpc.ltvmpc_kinetmatic_curvilinear(g["x0"], c_layout(g["x_ref"]), DT, c_layout(g["x_lin"]), c_layout(g["u_lin"]))
    _check_solution(r, g)


def test_fused_step_matches_golden_perturbed(mpc):
    g = load_golden("kinematic_perturbed_fsg2019.npz")
    r = mpc.ltvmpc_kinetmatic_curvilinear(g["x0"], c_layout(g["x_ref"]), DT, c_layout(g["x_lin"]), c_layout(g["u_lin"]))
    _check_solution(r, g)


def test_fused_step_other_track(mpc):
    g = load_golden("kinematic_lap_fso2020.npz")
    B = g["x0"].shape[0]
    r = mpc.ltvmpc_kinetmatic_curvilinear(g["x0"], c_layout(g["x_ref"]), DT, c_layout(g["x_lin"]), c_layout(g["u_lin"]),
                                          track_id=np.full(B, 2, np.int32))
    _check_solution(r, g)


def test_fused_step_live_oracle(mpc, fsg):
    """Fresh random perturbations solved by the oracle at test time (no fixture)."""
    from oracle import ltv
    g = load_golden("kinematic_lap_fsg2019.npz")
    rng = np.random.default_rng(123)
    B = 12
    pick = rng.integers(g["x0"].shape[0], size=B)
    x0 = g["x0"][pick].copy()
    x0[:, 1] += rng.uniform(-0.2, 0.2, B)
    x0[:, 3] = np.maximum(0.5, x0[:, 3] + rng.uniform(-1, 1, B))
    r = mpc.ltvmpc_kinetmatic_curvilinear(x0, c_layout(g["x_ref"][pick]), DT, c_layout(g["x_lin"][pick]),
                                          c_layout(g["u_lin"][pick]))
    for b in range(B):
        u, x, ef, fv, sl, sol = ltv.ltvmpc_kinetmatic_curvilinear(x0[b], g["x_ref"][pick[b]], fsg.kappa, DT,
                                                                   g["x_lin"][pick[b]], g["u_lin"][pick[b]])
        assert ef == r.exitflag[b] == 0
        assert rel(r.u_opt[b], u) <= U_RTOL
        assert rel(r.x_opt[b], x) <= U_RTOL


def test_batch_edge_cases(mpc):
    import fsae_mpc_b200 as fm
    g = load_golden("kinematic_lap_fsg2019.npz")
    # empty batch
    r = mpc.ltvmpc_kinetmatic_curvilinear(np.zeros((0, 5)), np.zeros((0, 40, 5)), DT, np.zeros((0, 40, 5)), np.zeros((0, 40, 2)))
    assert r.u_opt.shape == (0, 80)
    # single problem and a ragged (non multiple of anything) batch give identical per-problem results
    xr, xl, ul = c_layout(g["x_ref"]), c_layout(g["x_lin"]), c_layout(g["u_lin"])
    r1 = mpc.ltvmpc_kinetmatic_curvilinear(g["x0"][:1], xr[:1], DT, xl[:1], ul[:1])
    r7 = mpc.ltvmpc_kinetmatic_curvilinear(g["x0"][:7], xr[:7], DT, xl[:7], ul[:7])
    assert np.array_equal(r1.u_opt[0], r7.u_opt[0])
    # a horizon beyond the compiled capacity (FSAE_MAX_HORIZON = 80) is an error, not a silent fallback
    with pytest.raises(fm.FsaeError):
        mpc.ltvmpc_kinetmatic_curvilinear(g["x0"][:1], np.zeros((1, 81, 5)), DT, np.zeros((1, 81, 5)), np.zeros((1, 81, 2)))


def test_closed_loop_lap_matches_oracle_prefix(mpc, fsg):
    """main.m closed loop with the CUDA step in the loop vs the oracle step: the first 60
    steps of the lap must stay on the same trajectory."""
    from oracle import closed_loop as cl

    def gpu_step(x0, x_ref, kappa, dt, x_lin, u_lin):
        r = mpc.ltvmpc_kinetmatic_curvilinear(x0[None], c_layout(x_ref[None]), dt, c_layout(x_lin[None]), c_layout(u_lin[None]))
        return r.u_opt[0], r.x_opt[0], int(r.exitflag[0]), float(r.fval[0]), r.slack_opt[0]

    h_gpu = cl.run(fsg.track, "KINEMATIC", n_sim=60, mpc_step=gpu_step)
    h_ref = cl.run(fsg.track, "KINEMATIC", n_sim=60)
    assert h_gpu["steps"] == h_ref["steps"]
    assert np.max(np.abs(np.array(h_gpu["x"]) - np.array(h_ref["x"]))) < 1e-6
    assert all(e == 0 for e in h_gpu["exitflag"])


def test_kernel_variants_agree(mpc, mpc_x):
    """v2 (register-tiled product kernel, from the PRODUCT library) against v1 (shared-memory variant, which
    only the cross-check build of the library carries) on a larger, harder perturbed batch: same exit flags,
    same solutions."""
    from fsae_mpc_b200 import workload as wl
    x0, xr, xl, ul = wl.perturbed_batch("kinematic", "fsg2019", 2048, seed=7)
    r2 = mpc.ltvmpc_kinetmatic_curvilinear(x0, xr, DT, xl, ul)
    old = mpc_x.set_kernel_version(1)
    try:
        r1 = mpc_x.ltvmpc_kinetmatic_curvilinear(x0, xr, DT, xl, ul)
    finally:
        mpc_x.set_kernel_version(old)
    assert (r2.exitflag == 0).all() and (r1.exitflag == 0).all()
    scale = np.maximum(1.0, np.abs(r1.u_opt).max(axis=1))
    assert (np.abs(r2.u_opt - r1.u_opt).max(axis=1) / scale).max() < 1e-8
    assert np.abs(r2.x_opt - r1.x_opt).max() < 1e-7
    assert np.abs(r2.slack_opt - r1.slack_opt).max() < 1e-9
    same = (r1.workingSetB == r2.workingSetB).all(axis=1) & (r1.workingSetC == r2.workingSetC).all(axis=1)
    assert same.mean() > 0.99


@pytest.mark.parametrize("kv", [21, 26, 28, 29, 31, 32])
def test_warp_count_and_block_size_variants_agree(mpc, mpc_x, kv):
    """The dual active-set core is a template over the warp count (tile geometry) and the number of
    constraints taken per search (block size).  Every instantiation must reach the same minimiser and
    the same working set as the product configuration, whatever its pivot order.
    31 / 32: the solver in integrator coordinates w = T u (sparse normals), with the column-lane core and with the
    row-lane core (gi_core_rl.cuh) -- a different operator layout, coordinates and update formulas, same answers."""
    from fsae_mpc_b200 import workload as wl
    x0, xr, xl, ul = wl.perturbed_batch("kinematic", "fsg2019", 1500, seed=11)
    r2 = mpc.ltvmpc_kinetmatic_curvilinear(x0, xr, DT, xl, ul)
    old = mpc_x.set_kernel_version(kv)
    try:
        rv = mpc_x.ltvmpc_kinetmatic_curvilinear(x0, xr, DT, xl, ul)
    finally:
        mpc_x.set_kernel_version(old)
    assert (r2.exitflag == 0).all() and (rv.exitflag == 0).all()
    scale = np.maximum(1.0, np.abs(r2.u_opt).max(axis=1))
    assert (np.abs(rv.u_opt - r2.u_opt).max(axis=1) / scale).max() < 1e-8
    assert np.abs(rv.x_opt - r2.x_opt).max() < 1e-7
    assert (np.abs(rv.fval - r2.fval) / (1.0 + np.abs(r2.fval))).max() < 1e-10
    same = (rv.workingSetB == r2.workingSetB).all(axis=1) & (rv.workingSetC == r2.workingSetC).all(axis=1)
    assert same.mean() > 0.99


def test_sqp_passes_match_oracle_loop(mpc, fso):
    """BASELINE configs[3]: repeated relinearise+QP on fso2020, against the oracle iterated the
    same way (x_lin, u_lin <- previous x_opt, u_opt)."""
    import fsae_mpc_b200 as fm
    from oracle import ltv
    g = load_golden("kinematic_lap_fso2020.npz")
    pick = [2, 11, 25, 40]
    B = len(pick)
    n_sqp = 3
    r = mpc.ltvmpc_sqp(fm.KINEMATIC, g["x0"][pick], c_layout(g["x_ref"][pick]), DT, c_layout(g["x_lin"][pick]),
                       c_layout(g["u_lin"][pick]), n_sqp, track_id=np.full(B, 2, np.int32))
    assert (r.exitflag == 0).all()
    for j, b in enumerate(pick):
        xl, ul = g["x_lin"][b], g["u_lin"][b]
        for _ in range(n_sqp):
            u, x, ef, fv, sl, _s = ltv.ltvmpc_kinetmatic_curvilinear(g["x0"][b], g["x_ref"][b], fso.kappa, DT, xl, ul)
            assert ef == 0
            xl, ul = x.reshape(5, 40, order="F"), u.reshape(2, 40, order="F")
        assert rel(r.u_opt[j], u) <= 1e-6 and rel(r.x_opt[j], x) <= 1e-6
        assert abs(r.fval[j] - fv) <= 1e-6 * (1 + abs(fv))


@pytest.mark.parametrize("N", [20, 80])
def test_other_horizons_match_golden(mpc, N):
    """BASELINE configs[4]: horizons 20 and 80 (80: nV = 161, the operator lives in an
    L2-resident global slab instead of registers)."""
    g = load_golden(f"kinematic_lap_fsg2019_N{N}.npz")
    r = mpc.ltvmpc_kinetmatic_curvilinear(g["x0"], c_layout(g["x_ref"]), DT, c_layout(g["x_lin"]), c_layout(g["u_lin"]))
    _check_solution(r, g)


@pytest.mark.parametrize("model_name,N", [("kinematic", 40), ("kinematic", 20), ("kinematic", 80), ("dynamic", 40)])
def test_fused_kernel_taps_hessian_gradient_and_operator(mpc, model_name, N):
    """Stage-level check INSIDE the fused kernel: the Hessian it builds from the cost Gramian (one dot
    product per entry) and the gradient equal the condense stage's H / f (which the reference's own
    generate_qp.m pins, tests/test_reference_vectors.py), and the operator it gets from the Riccati
    recursion + adjoint rows (no factorisation) satisfies J'HJ = I with the slack columns leading."""
    import torch
    import fsae_mpc_b200 as fm
    model = fm.KINEMATIC if model_name == "kinematic" else fm.DYNAMIC
    fx = {("kinematic", 40): "kinematic_lap_fsg2019.npz", ("kinematic", 20): "kinematic_lap_fsg2019_N20.npz",
          ("kinematic", 80): "kinematic_lap_fsg2019_N80.npz", ("dynamic", 40): "dynamic_lap_fss2019.npz"}[(model_name, N)]
    g = load_golden(fx)
    B = min(6, g["x0"].shape[0])
    NX, NU, NS = (5, 2, 1) if model_name == "kinematic" else (7, 2, 4)
    nU, nV = NU * N, NU * N + NS
    tid = np.full(B, 0 if model_name == "kinematic" else 1, np.int32)
    pid = np.zeros(B, np.int32)
    if model_name == "dynamic":
        mpc.set_params(3, fm.default_params(fm.DYNAMIC))
        pid[:] = 3
    x0, xr, xl, ul = g["x0"][:B], c_layout(g["x_ref"][:B]), c_layout(g["x_lin"][:B]), c_layout(g["u_lin"][:B])
    ref = mpc.condense(model, x0, xr, DT, xl, ul, track_id=tid, param_id=pid)
    dev = torch.device("cuda", 0)
    d = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (x0, xr, xl, ul, tid, pid)]
    o = dict(u_opt=torch.empty((B, nU), dtype=torch.float64, device=dev), x_opt=torch.empty((B, NX * N), dtype=torch.float64, device=dev),
             exitflag=torch.empty(B, dtype=torch.int32, device=dev), fval=torch.empty(B, dtype=torch.float64, device=dev),
             slack_opt=torch.empty((B, NS), dtype=torch.float64, device=dev))
    tH = torch.zeros((B, nV, nV), dtype=torch.float64, device=dev)
    tg = torch.zeros((B, nV), dtype=torch.float64, device=dev)
    tM = torch.zeros((B, nV, nV), dtype=torch.float64, device=dev)
    ptrs = dict(x0=d[0].data_ptr(), x_ref=d[1].data_ptr(), x_lin=d[2].data_ptr(), u_lin=d[3].data_ptr(), track_id=d[4].data_ptr(),
                param_id=d[5].data_ptr(), **{k: v.data_ptr() for k, v in o.items()})
    mpc.set_taps(tH.data_ptr(), tg.data_ptr(), tM.data_ptr())
    try:
        mpc.ltvmpc_dev(model, B, N, DT, ptrs, stream=mpc.stream)
        torch.cuda.synchronize()
    finally:
        mpc.set_taps(0, 0, 0)
    H = tH.cpu().numpy().transpose(0, 2, 1)          # column-major per problem
    M = tM.cpu().numpy().transpose(0, 2, 1)
    gv = tg.cpu().numpy()
    Href = np.asarray(ref["H"]).reshape(B, nV, nV)
    fref = np.asarray(ref["f"]).reshape(B, nV)
    for b in range(B):
        hs = np.abs(Href[b]).max()
        assert np.abs(H[b][:nU, :nU] - Href[b][:nU, :nU]).max() <= 1e-11 * hs, (b, "H")
        assert np.abs(H[b] - H[b].T).max() == 0.0
        assert np.abs(gv[b] - fref[b]).max() <= 1e-10 * (1.0 + np.abs(fref[b]).max()), (b, "g")
        J = M[b][:nU, NS:]
        E = J.T @ Href[b][:nU, :nU] @ J - np.eye(nU)
        assert np.abs(E).max() <= 1e-9, (b, "J'HJ - I", np.abs(E).max())
        # slack columns lead: unit vectors on the slack rows, nothing else
        assert np.array_equal(M[b][:, :NS], np.eye(nV)[:, nU:])
        assert np.abs(M[b][nU:, NS:]).max() == 0.0


@pytest.mark.parametrize("model_name,N,B", [("kinematic", 40, 3000), ("kinematic", 20, 3000), ("kinematic", 80, 300), ("dynamic", 40, 1500)])
def test_results_are_deterministic(mpc, model_name, N, B):
    """The kernels synchronise warps through named barriers, double-buffered partial sums and a published stage
    flag; a missing fence or barrier would show up as run-to-run differences.  Same inputs, three runs (the batch
    is larger than the number of CTA slots, so block scheduling differs between runs): bit-identical outputs."""
    import fsae_mpc_b200 as fm
    from fsae_mpc_b200 import workload as wl
    if N == 80:
        g = load_golden("kinematic_lap_fsg2019_N80.npz")
        pick = np.arange(B) % g["x0"].shape[0]
        x0, xr, xl, ul = g["x0"][pick], c_layout(g["x_ref"])[pick], c_layout(g["x_lin"])[pick], c_layout(g["u_lin"])[pick]
        x0 = x0 + np.random.default_rng(1).uniform(-0.05, 0.05, x0.shape) * np.array([0, 1, 0.2, 1, 0.2])
    else:
        x0, xr, xl, ul = wl.perturbed_batch(model_name, "fsg2019" if model_name == "kinematic" else "fss2019", B, seed=21)
        xr, xl, ul = (np.ascontiguousarray(a[:, :N]) for a in (xr, xl, ul))
    kw = {}
    step = mpc.ltvmpc_kinetmatic_curvilinear
    if model_name == "dynamic":
        mpc.set_params(3, fm.default_params(fm.DYNAMIC))
        kw = dict(track_id=np.ones(B, np.int32), param_id=np.full(B, 3, np.int32))
        step = mpc.ltvmpc_dynamic_curvilinear
    runs = [step(x0, xr, DT, xl, ul, **kw) for _ in range(3)]
    for r in runs[1:]:
        for k in ("u_opt", "x_opt", "fval", "slack_opt", "exitflag", "iters", "workingSetB", "workingSetC"):
            assert np.array_equal(getattr(r, k), getattr(runs[0], k)), k
