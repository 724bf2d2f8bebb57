"""GPU: runtime horizon.  The reference computes N_steps = length(x_ref) at run time
(ltvmpc_kinetmatic_curvilinear.m:17, ltvmpc_dynamic_curvilinear.m:17); the kernels are compiled for the
capacities 20 / 40 / 80 and pad the remaining steps internally.  Any N in between must give the reference's
answer for THAT N: compared with the oracle step at the same (odd) horizon, outputs in the caller's shapes."""
import numpy as np
import pytest

from conftest import load_golden, c_layout, DT

pytestmark = pytest.mark.gpu

CASES = [("kinematic", "kinematic_lap_fsg2019.npz", "fsg2019", 0, 33), ("kinematic", "kinematic_lap_fsg2019.npz", "fsg2019", 0, 7),
         ("kinematic", "kinematic_lap_fsg2019.npz", "fsg2019", 0, 21), ("kinematic", "kinematic_lap_fsg2019_N80.npz", "fsg2019", 0, 55),
         ("kinematic", "kinematic_lap_fsg2019.npz", "fsg2019", 0, 5),     # the reference's minimum: N_steps = length(x_ref) =
                                                                          # max(size(x_ref)) is only the horizon when N_steps >= N_x
         ("dynamic", "dynamic_lap_fss2019.npz", "fss2019", 1, 27), ("dynamic", "dynamic_lap_fss2019.npz", "fss2019", 1, 13),
         # the dynamic model beyond horizon 40 (all packed B_bar rows in the L2 slab): the capacity itself and a padded one
         ("dynamic", "dynamic_lap_fss2019_N80.npz", "fss2019", 1, 80), ("dynamic", "dynamic_lap_fss2019_N80.npz", "fss2019", 1, 67)]


@pytest.mark.parametrize("model,fixture,track,tid,N", CASES)
def test_odd_horizons_match_oracle(mpc, model, fixture, track, tid, N):
    import fsae_mpc_b200 as fm
    from conftest import GoldenTrack
    from oracle import ltv
    g = load_golden(fixture)
    tr = GoldenTrack(track)
    mid = fm.KINEMATIC if model == "kinematic" else fm.DYNAMIC
    NX, NS, rows = (5, 1, 6) if model == "kinematic" else (7, 4, 20)
    pick = np.linspace(1, g["x0"].shape[0] - 1, 6).astype(int)
    x0 = g["x0"][pick]
    xr, xl, ul = (np.ascontiguousarray(g[k][pick][:, :, :N]) for k in ("x_ref", "x_lin", "u_lin"))    # first N steps
    pid = 40 + mid
    mpc.set_params(pid, fm.default_params(mid))
    step = mpc.ltvmpc_kinetmatic_curvilinear if model == "kinematic" else mpc.ltvmpc_dynamic_curvilinear
    r = step(x0, c_layout(xr), DT, c_layout(xl), c_layout(ul), track_id=np.full(len(pick), tid, np.int32),
             param_id=np.full(len(pick), pid, np.int32))
    assert r.u_opt.shape == (len(pick), 2 * N) and r.x_opt.shape == (len(pick), NX * N)
    assert r.workingSetB.shape == (len(pick), 2 * N + NS) and r.workingSetC.shape == (len(pick), rows * N)
    ostep = ltv.ltvmpc_kinetmatic_curvilinear if model == "kinematic" else ltv.ltvmpc_dynamic_curvilinear
    for j in range(len(pick)):
        u, x, ef, fv, sl, info = ostep(x0[j], xr[j], tr.kappa, DT, xl[j], ul[j])
        assert r.exitflag[j] == ef, (j, r.exitflag[j], ef)
        if ef != 0:
            continue
        assert np.max(np.abs(r.u_opt[j] - u)) <= 1e-6 * max(1.0, np.max(np.abs(u))), j          # north_star tolerance
        assert np.max(np.abs(r.x_opt[j] - x)) <= 1e-6 * max(1.0, np.max(np.abs(x))), j
        assert abs(r.fval[j] - fv) <= 1e-7 * (1 + abs(fv)) and np.max(np.abs(r.slack_opt[j] - sl)) <= 1e-7
        # same working set (north_star: "the same active set where the problem is non-degenerate").  At horizon 80 of
        # the dynamic model a few problems are degenerate -- several friction-polygon edges that share one slack are
        # active together and the multipliers are not unique -- so there up to four of the 1,764 entries may differ;
        # the minimiser is the same (checked above) and carries its own KKT certificate
        # (test_gpu_fullsize.py::test_dynamic_horizon_80_is_kkt_certified).
        ws_gpu = np.concatenate([r.workingSetB[j], r.workingSetC[j]])
        ws_ora = np.concatenate([info.workingSetB, info.workingSetC])
        assert (ws_gpu != ws_ora).sum() <= (4 if (model == "dynamic" and N > 40) else 0), (j, np.nonzero(ws_gpu != ws_ora)[0])

def test_horizon_limits_are_reported(mpc):
    import fsae_mpc_b200 as fm
    g = load_golden("kinematic_lap_fsg2019_N80.npz")
    x0 = g["x0"][:1]
    big = lambda a: np.ascontiguousarray(np.concatenate([a[:1], a[:1]], axis=2).transpose(0, 2, 1)[:, :81])
    with pytest.raises(fm.FsaeError, match="horizon"):
        mpc.ltvmpc_kinetmatic_curvilinear(x0, big(g["x_ref"]), DT, big(g["x_lin"]), big(g["u_lin"]))
