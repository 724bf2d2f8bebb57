"""Device-resident timing of the fused step (no host copies): python scripts/gpu_time_dev.py B [model track] ; env FSAE_KV selects the kernel variant."""
import os, sys, numpy as np, torch
sys.path.insert(0, '.')
import fsae_mpc_b200 as fm
from fsae_mpc_b200 import workload as wl
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
model = sys.argv[2] if len(sys.argv) > 2 else "kinematic"
track = sys.argv[3] if len(sys.argv) > 3 else "fsg2019"
mid = fm.KINEMATIC if model == "kinematic" else fm.DYNAMIC
NX, NU, NS = (5, 2, 1) if model == "kinematic" else (7, 2, 4)
N = 40
mpc = fm.FsaeMpc(0)
for tid, (n, t) in enumerate(wl.load_tracks().items()):
    mpc.set_track(tid, t[0], t[1], t[2])
if os.environ.get("FSAE_KV"): mpc.set_kernel_version(int(os.environ["FSAE_KV"]))
mpc.set_params(1, fm.default_params(mid))
x0, xr, xl, ul = wl.perturbed_batch(model, track, B, 0)
dev = torch.device("cuda", 0)
d = [torch.from_numpy(a).to(dev) for a in (x0, xr, xl, ul)]
tid_t = torch.full((B,), list(wl.load_tracks()).index(track), dtype=torch.int32, device=dev)
pid_t = torch.ones(B, dtype=torch.int32, device=dev)
o = dict(u_opt=torch.empty((B, NU * N), dtype=torch.float64, device=dev), x_opt=torch.empty((B, NX * N), dtype=torch.float64, device=dev),
         exitflag=torch.empty(B, dtype=torch.int32, device=dev), fval=torch.empty(B, dtype=torch.float64, device=dev),
         slack_opt=torch.empty((B, NS), dtype=torch.float64, device=dev), iters=torch.empty(B, dtype=torch.int32, device=dev))
ptrs = dict(x0=d[0].data_ptr(), x_ref=d[1].data_ptr(), x_lin=d[2].data_ptr(), u_lin=d[3].data_ptr(), track_id=tid_t.data_ptr(), param_id=pid_t.data_ptr(),
            **{k: v.data_ptr() for k, v in o.items()})
st = torch.cuda.ExternalStream(mpc.stream, device=dev)
for _ in range(3): mpc.ltvmpc_dev(mid, B, N, 0.05, ptrs, stream=mpc.stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 5
e0.record(st)
for _ in range(K): mpc.ltvmpc_dev(mid, B, N, 0.05, ptrs, stream=mpc.stream)
e1.record(st); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(f"{model} B={B} kv={os.environ.get('FSAE_KV','2')}: {ms:.2f} ms/step -> {B/ms*1e3:.0f} QP/s ; exit!=0 {(o['exitflag']!=0).sum().item()} iters {o['iters'].double().mean().item():.1f}")
if os.environ.get("FSAE_KV"):
    u_kv = o['u_opt'].clone(); f_kv = o['fval'].clone()
    mpc.set_kernel_version(1)            # shared-memory cross-check kernel
    mpc.ltvmpc_dev(mid, B, N, 0.05, ptrs, stream=mpc.stream); torch.cuda.synchronize()
    du = (u_kv - o['u_opt']).abs().max().item(); df = ((f_kv - o['fval']).abs() / (1 + o['fval'].abs())).max().item()
    print(f"   vs v1 kernel: max|du| {du:.3e}  max rel dfval {df:.3e}")
