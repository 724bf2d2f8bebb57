"""GPU, BASELINE.json's FULL batch sizes: every solution of the bench batches carries an optimality
certificate that does not involve the oracle.  For each problem the condense stage (pinned against the
reference's own generate_qp.m / *_state_constraints.m, tests/test_reference_vectors.py) supplies the QP
data H, f, xA, lb, ub, lbA, ubA; the fused step's answer z = [u_opt; slack_opt] and its working set must
satisfy the KKT conditions of THAT problem:

  primal feasibility   lb <= z <= ub, lbA <= xA z <= ubA
  active set           every constraint in the working set is at its bound
  stationarity         H z + f = C_act' lam        (lam by least squares on the working set)
  dual feasibility     lam >= 0 at lower bounds, <= 0 at upper bounds

A strictly convex QP (in u; exact penalty on the slacks) has ONE point with these properties: the one
qpOASES returns.  Infeasible problems (exitflag -2) are counted and must be rare.

The certificate's QP data come from the CUDA condense kernel, which shares its model / constraint code
(models.cuh, cons.cuh) with the fused kernel.  To take that common mode out, every certified batch also
compares the CUDA condense output of 256 randomly drawn problems OF THAT BATCH with the numpy oracle's
build_*_qp (itself pinned to the reference's .m files on 112 + 8 reference-executed problems):
H, f, xA, lb, ub, lbA, ubA to 1e-10 relative."""
import numpy as np
import pytest

from conftest import DT

pytestmark = pytest.mark.gpu
INF = 1e19


def _rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin)
    assert np.array_equal(np.sign(a[~fin]), np.sign(b[~fin]))
    return float(np.max(np.abs(a[fin] - b[fin])) / max(1.0, np.max(np.abs(b[fin])))) if fin.any() else 0.0


def _condense_matches_oracle(mpc, mid, model, track, x0, xr, xl, ul, t_id, p_id, n_check=256):
    """CUDA condense vs the oracle on n_check random problems of the batch (perturbed states included)."""
    from conftest import GoldenTrack
    from oracle import ltv
    tr = GoldenTrack(track)
    build = ltv.build_kinematic_qp if model == "kinematic" else ltv.build_dynamic_qp
    pick = np.sort(np.random.default_rng(7).choice(x0.shape[0], size=min(n_check, x0.shape[0]), replace=False))
    q = mpc.condense(mid, x0[pick], xr[pick], DT, xl[pick], ul[pick], track_id=t_id[pick], param_id=p_id[pick])
    worst = 0.0
    for j, b in enumerate(pick):
        o = build(x0[b], xr[b].T, tr.kappa, DT, xl[b].T, ul[b].T)
        for k in ("H", "f", "xA", "lb", "ub", "lbA", "ubA"):
            e = _rel(q[k][j], o[k])
            assert e < 1e-10, (k, int(b), e)
            worst = max(worst, e)
    return worst, len(pick)


def _certify(mpc, model, track, B, chunk, tid, pid, max_infeasible, N=40):
    import fsae_mpc_b200 as fm
    from fsae_mpc_b200 import workload as wl
    mid = fm.KINEMATIC if model == "kinematic" else fm.DYNAMIC
    if N == 80:                                                              # the committed horizon-80 laps, perturbed
        from conftest import load_golden, c_layout
        g = load_golden("kinematic_lap_fsg2019_N80.npz" if model == "kinematic" else "dynamic_lap_fss2019_N80.npz")
        rng = np.random.default_rng(1000)
        pick = rng.integers(g["x0"].shape[0], size=B)
        x0 = g["x0"][pick].copy()
        x0[:, 1] += rng.uniform(-0.3, 0.3, B); x0[:, 2] += rng.uniform(-0.08, 0.08, B)
        x0[:, 3] = np.maximum(0.5, x0[:, 3] + rng.uniform(-1.5, 1.5, B)); x0[:, -1] += rng.uniform(-0.05, 0.05, B)
        xr, xl, ul = c_layout(g["x_ref"])[pick], c_layout(g["x_lin"])[pick], c_layout(g["u_lin"])[pick]
    else:
        x0, xr, xl, ul = wl.perturbed_batch(model, track, B, seed=1000)      # bench.py's rank-0 batch
        xr, xl, ul = (np.ascontiguousarray(a[:, :N]) for a in (xr, xl, ul))  # N = 20: the first 20 steps
    step = mpc.ltvmpc_kinetmatic_curvilinear if model == "kinematic" else mpc.ltvmpc_dynamic_curvilinear
    t_id, p_id = np.full(B, tid, np.int32), np.full(B, pid, np.int32)
    r = step(x0, xr, DT, xl, ul, track_id=t_id, param_id=p_id)
    ok = r.exitflag == 0
    assert set(np.unique(r.exitflag)) <= {0, -2}
    assert (~ok).sum() <= max_infeasible, f"{(~ok).sum()} problems not solved"
    if model == "kinematic":
        # An infeasibility verdict must be one: the steering angle is a pure integrator with a hard bound,
        # delta_1 = delta_0 + dt * u_2, |u_2| <= 0.4, |delta_1| <= 0.4, so the problem is infeasible exactly when
        # the (perturbed) start value lies more than dt * 0.4 outside the bound.
        excess = np.abs(x0[:, 4]) - (0.4 + 0.4 * DT)
        clear = np.abs(excess) > 1e-7
        assert np.array_equal((~ok)[clear], (excess > 0)[clear]), "exitflag -2 does not coincide with infeasible steering"
    worst = dict(primal=0.0, active=0.0, stat=0.0, dual=0.0)
    worst["condense_vs_oracle"], worst["condense_checked"] = _condense_matches_oracle(mpc, mid, model, track, x0, xr, xl, ul, t_id, p_id)
    for lo in range(0, B, chunk):
        sl = slice(lo, min(B, lo + chunk))
        keep = ok[sl]
        if not keep.any():
            continue
        q = mpc.condense(mid, x0[sl], xr[sl], DT, xl[sl], ul[sl], track_id=t_id[sl], param_id=p_id[sl])
        H, f, A = q["H"][keep], q["f"][keep], q["xA"][keep]
        lb, ub, lbA, ubA = q["lb"][keep], q["ub"][keep], q["lbA"][keep], q["ubA"][keep]
        z = np.concatenate([r.u_opt[sl][keep], r.slack_opt[sl][keep]], axis=1)
        wsB, wsC = r.workingSetB[sl][keep].astype(np.int64), r.workingSetC[sl][keep].astype(np.int64)
        nb, n = z.shape
        Az = np.einsum("bmn,bn->bm", A, z)
        lo_all, up_all = np.concatenate([lb, lbA], 1), np.concatenate([ub, ubA], 1)
        val = np.concatenate([z, Az], 1)
        flo, fup = np.abs(lo_all) < INF, np.abs(up_all) < INF
        scale = 1.0 + np.minimum(np.where(flo, np.abs(lo_all), 0.0), np.where(fup, np.abs(up_all), 0.0))
        viol = np.maximum(np.where(flo, lo_all - val, -np.inf), np.where(fup, val - up_all, -np.inf)) / scale
        worst["primal"] = max(worst["primal"], float(viol.max()))
        ws = np.concatenate([wsB, wsC], 1)
        gap = np.where(ws < 0, np.abs(val - lo_all), np.where(ws > 0, np.abs(val - up_all), 0.0)) / scale
        worst["active"] = max(worst["active"], float(gap.max()))
        # multipliers by least squares on the working set (at most n constraints; padded with unit rows)
        grad = np.einsum("bij,bj->bi", H, z) + f
        order = np.argsort(ws == 0, axis=1, kind="stable")[:, :n]                # active constraints first
        act = np.take_along_axis(ws, order, 1) != 0
        assert (ws != 0).sum(1).max() <= n
        C = np.concatenate([np.broadcast_to(np.eye(n), (nb, n, n)), A], 1)
        Ca = np.take_along_axis(C, order[:, :, None], 1) * act[:, :, None]
        G = np.einsum("bqn,bpn->bqp", Ca, Ca) + np.eye(n)[None] * (~act)[:, :, None]
        lam = np.linalg.solve(G + 1e-30 * np.eye(n)[None], np.einsum("bqn,bn->bq", Ca, grad)[:, :, None])[:, :, 0]
        res = np.einsum("bqn,bq->bn", Ca, lam) - grad
        sstat = 1.0 + np.abs(grad).max(1) + np.abs(lam).max(1)
        worst["stat"] = max(worst["stat"], float((np.abs(res).max(1) / sstat).max()))
        side = np.take_along_axis(ws, order, 1)                                  # -1 lower: lam >= 0, +1 upper: lam <= 0
        curved = np.abs(np.einsum("bii->bi", H)) > 0
        dscale = 1.0 + np.where(curved, np.abs(grad), 0.0).max(1)
        dual = np.where(side < 0, -lam, np.where(side > 0, lam, 0.0)) / dscale[:, None]
        worst["dual"] = max(worst["dual"], float(dual.max()))
    return worst, int((~ok).sum())


def test_full_kinematic_batch_is_kkt_certified(mpc):
    """configs[1]: all 65,536 kinematic problems of the bench batch."""
    worst, ninf = _certify(mpc, "kinematic", "fsg2019", 65536, 4096, 0, 0, max_infeasible=0)
    print("kinematic 65,536: worst scaled KKT residuals", worst)
    assert worst["primal"] <= 1e-7 and worst["active"] <= 1e-7, worst
    assert worst["stat"] <= 1e-7 and worst["dual"] <= 1e-6, worst


def test_dynamic_horizon_80_is_kkt_certified(mpc):
    """configs[4] names horizons 20 / 40 / 80; the dynamic model (main.m:26, the reference's default) at 80: nV = 164,
    all packed B_bar rows, the full H and the J staging in the L2 slab.  1,024 perturbed problems of the committed
    horizon-80 lap on fss2019."""
    import fsae_mpc_b200 as fm
    mpc.set_params(3, fm.default_params(fm.DYNAMIC))
    worst, ninf = _certify(mpc, "dynamic", "fss2019", 1024, 256, 1, 3, max_infeasible=1024 // 20, N=80)
    print(f"dynamic N=80 fss2019 1,024: infeasible {ninf}, worst scaled KKT residuals", worst)
    assert worst["primal"] <= 1e-7 and worst["active"] <= 1e-7, worst
    assert worst["stat"] <= 1e-7 and worst["dual"] <= 1e-6, worst


def test_full_dynamic_batch_is_kkt_certified(mpc):
    """configs[2]: the 32,768 dynamic problems one GPU of the 8-GPU job solves."""
    import fsae_mpc_b200 as fm
    mpc.set_params(3, fm.default_params(fm.DYNAMIC))
    worst, ninf = _certify(mpc, "dynamic", "fss2019", 32768, 1024, 1, 3, max_infeasible=0)
    print("dynamic 32,768: worst scaled KKT residuals", worst)
    assert worst["primal"] <= 1e-7 and worst["active"] <= 1e-7, worst
    assert worst["stat"] <= 1e-7 and worst["dual"] <= 1e-6, worst


@pytest.mark.parametrize("N,track,tid,B,chunk", [(20, "fso2020", 2, 16384, 8192), (80, "fsg2019", 0, 2048, 512), (40, "fss2019", 1, 8192, 4096)])
def test_sweep_batches_are_kkt_certified(mpc, N, track, tid, B, chunk):
    """configs[4]: other horizons and tracks (horizon 80 runs the split register/shared-memory tile).  Perturbed
    start states can make a problem infeasible (steering beyond its hard bound at step 0): those are reported as
    exitflag -2, everything else must be a KKT point."""
    worst, ninf = _certify(mpc, "kinematic", track, B, chunk, tid, 0, max_infeasible=B // 20, N=N)
    print(f"kinematic N={N} {track} {B}: infeasible {ninf}, worst scaled KKT residuals", worst)
    assert worst["primal"] <= 1e-7 and worst["active"] <= 1e-7, worst
    assert worst["stat"] <= 1e-7 and worst["dual"] <= 1e-6, worst
