"""End-to-end rate of fsae_ltvmpc_host from PAGEABLE host buffers (what a MEX gateway passes) against the
device-resident rate, for the three workloads VERDICT r1 lists.   python scripts/bench_pageable.py [steps]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import fsae_mpc_b200 as fm
from fsae_mpc_b200 import workload as wl
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
mpc = fm.FsaeMpc(0)
for tid, (n, t) in enumerate(wl.load_tracks().items()):
    mpc.set_track(tid, t[0], t[1], t[2])
mpc.set_params(1, fm.default_params(fm.DYNAMIC))
for name, model, track, tid, pid, B, N in (("kinematic N=40", "kinematic", "fsg2019", 0, 0, 65536, 40),
                                           ("kinematic N=20", "kinematic", "fsg2019", 0, 0, 65536, 20),
                                           ("dynamic N=40", "dynamic", "fss2019", 1, 1, 32768, 40)):
    x0, xr, xl, ul = wl.perturbed_batch(model, track, B, seed=1000)
    xr, xl, ul = (np.ascontiguousarray(a[:, :N]) for a in (xr, xl, ul))
    step = mpc.ltvmpc_kinetmatic_curvilinear if model == "kinematic" else mpc.ltvmpc_dynamic_curvilinear
    ids = dict(track_id=np.full(B, tid, np.int32), param_id=np.full(B, pid, np.int32))
    res = {}
    # the same call through the C-ABI with caller-owned, REUSED output arrays (no first-touch page faults per call)
    import ctypes as C
    NXm, NUm, NSm = (5, 2, 1) if model == "kinematic" else (7, 2, 4)
    outs = [np.empty((B, NUm * N)), np.empty((B, NXm * N)), np.empty(B, np.int32), np.empty(B), np.empty((B, NSm))]
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double)); ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    midm = fm.KINEMATIC if model == "kinematic" else fm.DYNAMIC
    def raw():
        rc = mpc._lib.fsae_ltvmpc_host(mpc._ctx, midm, B, N, 0.05, ip(ids["track_id"]), ip(ids["param_id"]), dp(x0), dp(xr), dp(xl), dp(ul),
                                       dp(outs[0]), dp(outs[1]), ip(outs[2]), dp(outs[3]), dp(outs[4]), None, None, None)
        assert rc == 0
    mpc.set_host_staging(0)
    for _ in range(2): raw()
    t0 = time.perf_counter()
    for _ in range(steps): raw()
    res["ring_reused_out"] = B * steps / (time.perf_counter() - t0)
    for mode, label in ((1, "direct"), (0, "ring")):
        mpc.set_host_staging(mode)
        for _ in range(2):
            step(x0, xr, DT := 0.05, xl, ul, **ids)
        t0 = time.perf_counter()
        for _ in range(steps):
            step(x0, xr, 0.05, xl, ul, **ids)
        res[label] = B * steps / (time.perf_counter() - t0)
    # device-resident rate
    dev = torch.device("cuda", 0)
    mid = fm.KINEMATIC if model == "kinematic" else fm.DYNAMIC
    NX, NU, NS = (5, 2, 1) if model == "kinematic" else (7, 2, 4)
    d = [torch.from_numpy(a).to(dev) for a in (x0, xr, xl, ul)]
    o = dict(u_opt=torch.empty((B, NU * N), dtype=torch.float64, device=dev), x_opt=torch.empty((B, NX * N), dtype=torch.float64, device=dev),
             exitflag=torch.empty(B, dtype=torch.int32, device=dev), fval=torch.empty(B, dtype=torch.float64, device=dev),
             slack_opt=torch.empty((B, NS), dtype=torch.float64, device=dev))
    t_ = torch.from_numpy(ids["track_id"]).to(dev); p_ = torch.from_numpy(ids["param_id"]).to(dev)
    ptrs = dict(x0=d[0].data_ptr(), x_ref=d[1].data_ptr(), x_lin=d[2].data_ptr(), u_lin=d[3].data_ptr(), track_id=t_.data_ptr(), param_id=p_.data_ptr(),
                **{k: v.data_ptr() for k, v in o.items()})
    for _ in range(2):
        mpc.ltvmpc_dev(mid, B, N, 0.05, ptrs, stream=mpc.stream)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        mpc.ltvmpc_dev(mid, B, N, 0.05, ptrs, stream=mpc.stream)
    torch.cuda.synchronize()
    res["device"] = B * steps / (time.perf_counter() - t0)
    print(json.dumps({"workload": name, "batch": B, "device_qps": res["device"], "pageable_direct_qps": res["direct"],
                      "pageable_ring_qps": res["ring"], "pageable_ring_reused_outputs_qps": res["ring_reused_out"], "ring_over_device": res["ring"] / res["device"],
                      "direct_over_device": res["direct"] / res["device"], "copy_threads": os.environ.get("FSAE_COPY_THREADS", "default")}), flush=True)
