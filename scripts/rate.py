"""Device-resident rate + active-set event counts of the fused step for the named workloads (A/B of kernel builds:
FSAE_LIB=build/libfsae_<variant>.so python scripts/rate.py [workloads...]).  Workloads: kin40 kin20 kin80 dyn40 dyn20."""
import json, os, sys, numpy as np, torch
sys.path.insert(0, '.')
import fsae_mpc_b200 as fm
from fsae_mpc_b200 import workload as wl
W = {"kin40": ("kinematic", "fsg2019", 65536, 40), "kin20": ("kinematic", "fsg2019", 65536, 20), "kin80": ("kinematic", "fsg2019", 4096, 80),
     "dyn40": ("dynamic", "fss2019", 32768, 40), "dyn20": ("dynamic", "fss2019", 32768, 20)}
names = [a for a in sys.argv[1:] if a in W] or ["kin40", "kin20", "dyn40"]
mpc = fm.FsaeMpc(0)
if os.environ.get("FSAE_KV"):            # kernel variant of the cross-check library (FSAE_LIB=fsae_mpc_b200/libfsae_mpc_b200_xcheck.so)
    mpc.set_kernel_version(int(os.environ["FSAE_KV"]))
tracks = wl.load_tracks()
for tid, (n, t) in enumerate(tracks.items()):
    mpc.set_track(tid, t[0], t[1], t[2])
mpc.set_params(1, fm.default_params(fm.DYNAMIC))
dev = torch.device("cuda", 0)
st = torch.cuda.ExternalStream(mpc.stream, device=dev)
for name in names:
    model, track, B, N = W[name]
    mid = fm.KINEMATIC if model == "kinematic" else fm.DYNAMIC
    NX, NU, NS = (5, 2, 1) if model == "kinematic" else (7, 2, 4)
    if N == 80:
        g = dict(np.load(os.path.join(wl.GOLDEN, "kinematic_lap_fsg2019_N80.npz")))
        rng = np.random.default_rng(1000)
        pick = rng.integers(g["x0"].shape[0], size=B)
        x0 = g["x0"][pick].copy()
        x0[:, 1] += rng.uniform(-0.3, 0.3, B); x0[:, 2] += rng.uniform(-0.08, 0.08, B)
        tr_ = lambda a: np.ascontiguousarray(a.transpose(0, 2, 1)[pick])
        xr, xl, ul = tr_(g["x_ref"]), tr_(g["x_lin"]), tr_(g["u_lin"])
    else:
        x0, xr, xl, ul = wl.perturbed_batch(model, track, B, 1000)
        xr, xl, ul = (np.ascontiguousarray(a[:, :N]) for a in (xr, xl, ul))
    d = [torch.from_numpy(a).to(dev) for a in (x0, xr, xl, ul)]
    tid_t = torch.full((B,), list(tracks).index(track), dtype=torch.int32, device=dev)
    pid_t = torch.full((B,), 0 if model == "kinematic" else 1, dtype=torch.int32, device=dev)
    o = dict(u_opt=torch.empty((B, NU * N), dtype=torch.float64, device=dev), x_opt=torch.empty((B, NX * N), dtype=torch.float64, device=dev),
             exitflag=torch.empty(B, dtype=torch.int32, device=dev), fval=torch.empty(B, dtype=torch.float64, device=dev),
             slack_opt=torch.empty((B, NS), dtype=torch.float64, device=dev), iters=torch.empty(B, dtype=torch.int32, device=dev))
    ptrs = dict(x0=d[0].data_ptr(), x_ref=d[1].data_ptr(), x_lin=d[2].data_ptr(), u_lin=d[3].data_ptr(), track_id=tid_t.data_ptr(), param_id=pid_t.data_ptr(),
                **{k: v.data_ptr() for k, v in o.items()})
    for _ in range(3): mpc.ltvmpc_dev(mid, B, N, 0.05, ptrs, stream=mpc.stream)
    torch.cuda.synchronize()
    mpc.counters(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 5
    e0.record(st)
    for _ in range(K): mpc.ltvmpc_dev(mid, B, N, 0.05, ptrs, stream=mpc.stream)
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    a, dr, rf = mpc.counters()
    print(json.dumps({"workload": name, "lib": os.path.basename(os.environ.get("FSAE_LIB", "product")) + (":kv" + os.environ["FSAE_KV"] if os.environ.get("FSAE_KV") else ""), "batch": B, "ms_per_step": round(ms, 3),
                      "qps": round(B / ms * 1e3), "exit_nonzero": int((o['exitflag'] != 0).sum().item()),
                      "iters": round(o['iters'].double().mean().item(), 2), "adds": round(a / (B * K), 2), "drops": round(dr / (B * K), 2),
                      "refreshes": round(rf / (B * K), 3), "checksum_u": float(o['u_opt'].double().abs().sum().item())}), flush=True)
