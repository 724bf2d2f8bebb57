"""Device-resident timing of the horizon-20 kernel variants (warp count x CTAs per SM)."""
import os, sys, numpy as np, torch
sys.path.insert(0, '.')
import fsae_mpc_b200 as fm
from fsae_mpc_b200 import workload as wl
B = 65536; N = 20; NX, NU, NS = 5, 2, 1
mpc = fm.FsaeMpc(0)
for tid, (n, t) in enumerate(wl.load_tracks().items()): mpc.set_track(tid, t[0], t[1], t[2])
x0, xr, xl, ul = wl.perturbed_batch("kinematic", "fsg2019", B, 0)
c = lambda a: np.ascontiguousarray(a[:, :N])
dev = torch.device("cuda", 0)
d = [torch.from_numpy(a).to(dev) for a in (x0, c(xr), c(xl), c(ul))]
o = dict(u_opt=torch.empty((B, NU * N), dtype=torch.float64, device=dev), x_opt=torch.empty((B, NX * N), dtype=torch.float64, device=dev),
         exitflag=torch.empty(B, dtype=torch.int32, device=dev), fval=torch.empty(B, dtype=torch.float64, device=dev),
         slack_opt=torch.empty((B, NS), dtype=torch.float64, device=dev), iters=torch.empty(B, dtype=torch.int32, device=dev))
ptrs = dict(x0=d[0].data_ptr(), x_ref=d[1].data_ptr(), x_lin=d[2].data_ptr(), u_lin=d[3].data_ptr(), **{k: v.data_ptr() for k, v in o.items()})
st = torch.cuda.ExternalStream(mpc.stream, device=dev)
ref = None
for kv in (2, 21, 26, 28, 29):
    mpc.set_kernel_version(kv)
    for _ in range(2): mpc.ltvmpc_dev(fm.KINEMATIC, B, N, 0.05, ptrs, stream=mpc.stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(5): mpc.ltvmpc_dev(fm.KINEMATIC, B, N, 0.05, ptrs, stream=mpc.stream)
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    u = o["u_opt"].clone()
    if ref is None: ref = u
    print(f"kv={kv}: {ms:.2f} ms -> {B/ms*1e3:.0f} QP/s ; max|du| vs default {(u-ref).abs().max().item():.2e} ; exit!=0 {(o['exitflag']!=0).sum().item()}")
