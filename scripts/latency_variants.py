"""Single-problem latency (one MPC step of one vehicle, device-resident) of the kernel variants in the cross-check
library: which operator layout / warp count has the shortest critical path when a CTA has an SM to itself.
    FSAE_LIB=fsae_mpc_b200/libfsae_mpc_b200_xcheck.so python scripts/latency_variants.py"""
import json, os, sys, numpy as np, torch
sys.path.insert(0, '.')
import fsae_mpc_b200 as fm
from fsae_mpc_b200 import workload as wl
mpc = fm.FsaeMpc(0)
tracks = wl.load_tracks()
for tid, (n, t) in enumerate(tracks.items()):
    mpc.set_track(tid, t[0], t[1], t[2])
dev = torch.device("cuda", 0)
st = torch.cuda.ExternalStream(mpc.stream, device=dev)
for N in (40, 20):
    for B in (1, 16, 148):
        x0, xr, xl, ul = wl.perturbed_batch("kinematic", "fsg2019", 256, 1000)
        xr, xl, ul = (np.ascontiguousarray(a[:, :N]) for a in (xr, xl, ul))
        d = [torch.from_numpy(a).to(dev) for a in (x0, xr, xl, ul)]
        o = dict(u_opt=torch.empty((256, 2 * N), dtype=torch.float64, device=dev), x_opt=torch.empty((256, 5 * N), dtype=torch.float64, device=dev),
                 exitflag=torch.empty(256, dtype=torch.int32, device=dev), fval=torch.empty(256, dtype=torch.float64, device=dev),
                 slack_opt=torch.empty((256, 1), dtype=torch.float64, device=dev), iters=torch.empty(256, dtype=torch.int32, device=dev))
        for kv in (2, 21, 22, 23, 24, 25, 29, 31, 32) if N == 40 else (2, 21, 26, 31, 32):
            try:
                mpc.set_kernel_version(kv)
            except Exception as e:
                continue
            lat = []
            for rep in range(64):
                i0 = (rep * 3) % (256 - B + 1)
                ptrs = dict(x0=d[0][i0:].data_ptr(), x_ref=d[1][i0:].data_ptr(), x_lin=d[2][i0:].data_ptr(), u_lin=d[3][i0:].data_ptr(),
                            **{k: v.data_ptr() for k, v in o.items()})
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                mpc.ltvmpc_dev(fm.KINEMATIC, B, N, 0.05, ptrs, stream=mpc.stream)
                e1.record(st); torch.cuda.synchronize()
                if rep >= 8: lat.append(e0.elapsed_time(e1) * 1e3)
            print(json.dumps({"N": N, "B": B, "kv": kv, "us_p50": round(float(np.percentile(lat, 50)), 1), "us_p90": round(float(np.percentile(lat, 90)), 1),
                              "iters": round(o["iters"][:B].double().mean().item(), 1)}), flush=True)
