"""BASELINE.json configs[4]: "scaling sweep: 1M QPs across horizons 20/40/80 and all three tracks at 1/2/4/8 GPUs vs
host-core qpOASES".  A FIXED job of ~1.02 M QPs (strong scaling): every case's batch is split over the ranks with
fsae_shard_range, each rank solves its shard device-resident, the case time is the MAX over ranks (the only
collective: one all-reduce of the timing vector at the end, north_star "only a final scalar or metric gather").
    python scripts/bench_sweep.py                       # one GPU
    torchrun --nproc-per-node 8 scripts/bench_sweep.py  # 8 GPUs, same total job
Rank 0 also times the C port of the reference path (oracle/ltvmpc_oracle.c, kind "port") on the host cores on a
bounded sample of every case and prints ONE JSON line."""
import json, os, sys, time
import numpy as np, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fsae_mpc_b200 as fm
from fsae_mpc_b200 import workload as wl
from fsae_mpc_b200.sharding import shard_range

rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
scale = float(os.environ.get("FSAE_SWEEP_SCALE", "1"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
mpc = fm.FsaeMpc(local)
tracks = wl.load_tracks()
tnames = list(tracks)
for tid, n in enumerate(tnames):
    mpc.set_track(tid, *tracks[n][:3])
mpc.set_params(1, fm.default_params(fm.DYNAMIC))
st = torch.cuda.ExternalStream(mpc.stream, device=dev)
DT = 0.05
# (model, track, horizon, global batch, CPU sample)
CASES = [("kinematic", t, 20, 131072, 2048) for t in tnames] + [("kinematic", t, 40, 131072, 1024) for t in tnames] + \
        [("kinematic", "fsg2019", 80, 32768, 128), ("dynamic", "fss2019", 20, 65536, 512), ("dynamic", "fss2019", 40, 131072, 256),
         ("dynamic", "fss2019", 80, 8192, 32)]
STEPS = 2
res, times = [], []
for ci, (model, track, N, Bg, cpu_n) in enumerate(CASES):
    Bg = max(world, int(Bg * scale))
    lo, hi = shard_range(Bg, rank, world)
    B = hi - lo
    mid = fm.KINEMATIC if model == "kinematic" else fm.DYNAMIC
    NX, NU, NS = (5, 2, 1) if model == "kinematic" else (7, 2, 4)
    x0, xr, xl, ul, track = wl.horizon_batch(model, track, B, N, seed=5000 + 97 * ci + rank)
    d = [torch.from_numpy(a).to(dev) for a in (x0, xr, xl, ul)]
    tid_t = torch.full((B,), tnames.index(track), dtype=torch.int32, device=dev)
    pid_t = torch.full((B,), 1 if model == "dynamic" else 0, dtype=torch.int32, device=dev)
    o = dict(u_opt=torch.empty((B, NU * N), dtype=torch.float64, device=dev), x_opt=torch.empty((B, NX * N), dtype=torch.float64, device=dev),
             exitflag=torch.empty(B, dtype=torch.int32, device=dev), fval=torch.empty(B, dtype=torch.float64, device=dev),
             slack_opt=torch.empty((B, NS), dtype=torch.float64, device=dev), iters=torch.empty(B, dtype=torch.int32, device=dev))
    ptrs = dict(x0=d[0].data_ptr(), x_ref=d[1].data_ptr(), x_lin=d[2].data_ptr(), u_lin=d[3].data_ptr(), track_id=tid_t.data_ptr(),
                param_id=pid_t.data_ptr(), **{k: v.data_ptr() for k, v in o.items()})
    mpc.ltvmpc_dev(mid, B, N, DT, ptrs, stream=mpc.stream)                    # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(STEPS):
        mpc.ltvmpc_dev(mid, B, N, DT, ptrs, stream=mpc.stream)
    e1.record(st); torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1) / STEPS)
    ef = o["exitflag"]
    res.append(dict(model=model, track=track, horizon=N, global_batch=Bg, infeasible=int((ef == -2).sum().item()),
                    max_iter=int((ef == 1).sum().item()), other_nonzero=int(((ef != 0) & (ef != -2) & (ef != 1)).sum().item()),
                    iters_sum=float(o["iters"].double().sum().item())))
    if rank == 0:
        res[-1]["_cpu_in"] = (x0[:cpu_n], xr[:cpu_n], xl[:cpu_n], ul[:cpu_n])
    del d, o
    torch.cuda.empty_cache()
t = torch.tensor(times, dtype=torch.float64, device=dev)
cnt = torch.tensor([[r["infeasible"], r["max_iter"], r["other_nonzero"], r["iters_sum"]] for r in res], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)          # the job's time per case = the slowest rank
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
if rank == 0:
    import cpu_baseline
    cases = []
    for i, r in enumerate(res):
        x0, xr, xl, ul = r.pop("_cpu_in")
        base = (cpu_baseline.Baseline if r["model"] == "kinematic" else cpu_baseline.DynamicBaseline)(tracks[r["track"]])
        n = x0.shape[0]
        base.run(x0[: max(1, n // 8)], xr[: max(1, n // 8)], xl[: max(1, n // 8)], ul[: max(1, n // 8)], DT)
        c0 = time.perf_counter(); base.run(x0, xr, xl, ul, DT); c1 = time.perf_counter()
        ms = float(t[i].item())
        cases.append(dict(r, ms_per_pass=ms, qp_per_s=r["global_batch"] / ms * 1e3, infeasible=int(cnt[i, 0].item()),
                          max_iter=int(cnt[i, 1].item()), other_nonzero=int(cnt[i, 2].item()),
                          iters_mean=float(cnt[i, 3].item()) / r["global_batch"],
                          cpu_port_qp_per_s=n / (c1 - c0), cpu_cores=base.cores, cpu_sample=n))
        cases[-1].pop("iters_sum")
    tot = sum(c["global_batch"] for c in cases); tms = sum(c["ms_per_pass"] for c in cases)
    cpu_s = sum(c["global_batch"] / c["cpu_port_qp_per_s"] for c in cases)
    print(json.dumps({"config": "configs[4] sweep: horizons 20/40/80 x tracks, kinematic + dynamic, fixed job (strong scaling)",
                      "n_gpus": world, "total_qp": tot, "total_ms": tms, "aggregate_qp_per_s": tot / tms * 1e3,
                      "cpu_port_extrapolated_s": cpu_s, "cpu_port_aggregate_qp_per_s": tot / cpu_s,
                      "cpu_kind": "port (oracle/ltvmpc_oracle.c on all host cores, bounded sample per case)", "cases": cases}))
if world > 1:
    dist.destroy_process_group()
