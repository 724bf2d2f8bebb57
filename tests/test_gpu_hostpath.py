"""GPU: the two host paths of fsae_ltvmpc_host -- direct copies (pinned callers) and the pinned
staging ring with copy threads (pageable callers: MATLAB mxArrays, plain numpy) -- return bit-identical
results, on ragged batch sizes that do not divide into the ring's chunks."""
import numpy as np
import pytest

from conftest import DT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("model,track,tid,B", [("kinematic", "fsg2019", 0, 10000), ("kinematic", "fsg2019", 0, 2049),
                                                ("dynamic", "fss2019", 1, 9001)])
def test_staging_ring_equals_direct_copies(mpc, model, track, tid, B):
    import fsae_mpc_b200 as fm
    from fsae_mpc_b200 import workload as wl
    mid = fm.KINEMATIC if model == "kinematic" else fm.DYNAMIC
    pid = 30 + mid
    mpc.set_params(pid, fm.default_params(mid))
    x0, xr, xl, ul = wl.perturbed_batch(model, track, B, seed=77)        # plain numpy: pageable
    ids = dict(track_id=np.full(B, tid, np.int32), param_id=np.full(B, pid, np.int32))
    step = mpc.ltvmpc_kinetmatic_curvilinear if model == "kinematic" else mpc.ltvmpc_dynamic_curvilinear
    try:
        mpc.set_host_staging(1)
        a = step(x0, xr, DT, xl, ul, **ids)
        assert mpc.last_host_path == 0
        mpc.set_host_staging(0)                                          # automatic: pageable -> ring
        b = step(x0, xr, DT, xl, ul, **ids)
        assert mpc.last_host_path == 1
        b2 = step(x0, xr, DT, xl, ul, **ids)                             # ring slots reused
    finally:
        mpc.set_host_staging(0)
    for k in ("u_opt", "x_opt", "exitflag", "fval", "slack_opt", "iters", "workingSetB", "workingSetC"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
        assert np.array_equal(getattr(a, k), getattr(b2, k)), k
    assert (a.exitflag == 0).mean() > 0.9


def test_small_batches_copy_directly(mpc):
    import fsae_mpc_b200 as fm
    from fsae_mpc_b200 import workload as wl
    x0, xr, xl, ul = wl.perturbed_batch("kinematic", "fsg2019", 64, seed=3)
    r = mpc.ltvmpc_kinetmatic_curvilinear(x0, xr, DT, xl, ul)
    assert mpc.last_host_path == 0 and (r.exitflag == 0).all()


@pytest.mark.parametrize("n_ctx", [1, 2, 3])
def test_device_pool_equals_single_context(mpc, n_ctx):
    """fsae_ltvmpc_host_pool: one host thread, n contexts (here all on device 0 -- the split and the
    concurrency are what is tested; profiles/ holds the 2- and 8-GPU runs): bit-identical to one context."""
    import fsae_mpc_b200 as fm
    from fsae_mpc_b200 import workload as wl
    from conftest import load_golden
    B = 5003
    x0, xr, xl, ul = wl.perturbed_batch("kinematic", "fsg2019", B, seed=5)
    ref = mpc.ltvmpc_kinetmatic_curvilinear(x0, xr, DT, xl, ul)
    pool = fm.FsaePool(devices=[0] * n_ctx)
    try:
        assert len(pool) == n_ctx
        t = load_golden("tracks.npz")
        pool.set_track(0, t["fsg2019_x"], t["fsg2019_y"], float(t["fsg2019_dl"]))
        assert sum(h - l for l, h in (pool.shard_range(B, r) for r in range(n_ctx))) == B
        r = pool.ltvmpc(fm.KINEMATIC, x0, xr, DT, xl, ul)
    finally:
        pool.close()
    for k in ("u_opt", "x_opt", "exitflag", "fval", "slack_opt", "iters", "workingSetB", "workingSetC"):
        assert np.array_equal(getattr(ref, k), getattr(r, k)), k
