"""ONE host process drives every visible GPU through the device pool (fsae_pool_create / fsae_ltvmpc_host_pool) -- the
path a single MATLAB process takes (matlab/fsae_mpc_b200_handle.m).  Pageable numpy buffers in and out, wall clock
around the C-ABI call: the number a MEX caller would see.  Weak scaling: 65,536 kinematic problems per GPU.
    python scripts/bench_pool.py [n_devices]"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fsae_mpc_b200 as fm
from fsae_mpc_b200 import workload as wl
nvis = torch.cuda.device_count()
n = int(sys.argv[1]) if len(sys.argv) > 1 else nvis
per = int(os.environ.get("FSAE_POOL_BATCH", 65536))
tracks = wl.load_tracks()
t = tracks["fsg2019"]
out = []
for model, mid, trk, per_gpu in (("kinematic", fm.KINEMATIC, "fsg2019", per), ("dynamic", fm.DYNAMIC, "fss2019", per // 2)):
    pool = fm.FsaePool(devices=list(range(n)))
    for tid, name in enumerate(tracks):
        pool.set_track(tid, *tracks[name][:3])
    pool.set_params(1, fm.default_params(fm.DYNAMIC))
    B = per_gpu * n
    x0, xr, xl, ul = wl.perturbed_batch(model, trk, B, seed=1000)
    tid = np.full(B, list(tracks).index(trk), np.int32); pid = np.full(B, 1 if model == "dynamic" else 0, np.int32)
    r = pool.ltvmpc(mid, x0, xr, 0.05, xl, ul, track_id=tid, param_id=pid)            # warm-up (allocations, staging rings)
    ts = []
    for _ in range(4):
        t0 = time.perf_counter(); r = pool.ltvmpc(mid, x0, xr, 0.05, xl, ul, track_id=tid, param_id=pid); ts.append(time.perf_counter() - t0)
    # the single-context answer for the first and the last shard's first problems: the split must not change results
    one = fm.FsaeMpc(0)
    for i, name in enumerate(tracks):
        one.set_track(i, *tracks[name][:3])
    one.set_params(1, fm.default_params(fm.DYNAMIC))
    step = one.ltvmpc_kinetmatic_curvilinear if model == "kinematic" else one.ltvmpc_dynamic_curvilinear
    sel = np.r_[0:256, B - 256:B]
    r1 = step(x0[sel], xr[sel], 0.05, xl[sel], ul[sel], track_id=tid[sel], param_id=pid[sel])
    same = bool(np.array_equal(r1.u_opt, r.u_opt[sel]) and np.array_equal(r1.exitflag, r.exitflag[sel]))
    one.close()
    best = min(ts)
    out.append({"model": model, "n_devices": n, "batch": B, "wall_ms_best": best * 1e3, "wall_ms_all": [round(x * 1e3, 2) for x in ts],
                "e2e_pageable_qp_per_s": B / best, "exitflag_nonzero": int((r.exitflag != 0).sum()),
                "bit_identical_to_single_context_on_512_problems": same})
    pool.close()
print(json.dumps({"api": "fsae_ltvmpc_host_pool (one host thread, pageable buffers, all devices)", "results": out}))
