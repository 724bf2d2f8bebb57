"""The other BASELINE.json configs, one JSON line each (device-resident kernel time with CUDA events on the
context's stream; `e2e` through the host-buffer API where stated):
  configs[2]  dynamic model, fss2019, 32,768 problems per GPU (262,144 over 8 GPUs)
  configs[3]  SQP: 3 relinearise+QP passes, batch 16,384, fso2020
  configs[4]  sweep over horizons 20 / 40 / 80 and the three tracks
    python scripts/bench_configs.py [--quick]
Under torchrun every rank runs its own shard (no collective); rank 0 prints its own numbers."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fsae_mpc_b200 as fm
from fsae_mpc_b200 import workload as wl

quick = "--quick" in sys.argv
rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
mpc = fm.FsaeMpc(local)
tracks = wl.load_tracks()
tnames = list(tracks)
for tid, n in enumerate(tnames):
    mpc.set_track(tid, *tracks[n][:3])
mpc.set_params(1, fm.default_params(fm.DYNAMIC))
st = torch.cuda.ExternalStream(mpc.stream, device=dev)
DT = 0.05


def horizon_batch(model, track, B, N, seed):
    """N = 40: lap samples; N = 20: the first 20 steps of the same problems; N = 80: the committed fsg2019 lap."""
    if N == 80:
        g = dict(np.load(os.path.join(wl.GOLDEN, "kinematic_lap_fsg2019_N80.npz")))
        rng = np.random.default_rng(seed)
        pick = rng.integers(g["x0"].shape[0], size=B)
        x0 = g["x0"][pick].copy()
        x0[:, 1] += rng.uniform(-0.3, 0.3, B); x0[:, 2] += rng.uniform(-0.08, 0.08, B)
        x0[:, 3] = np.maximum(0.5, x0[:, 3] + rng.uniform(-1.5, 1.5, B)); x0[:, -1] += rng.uniform(-0.05, 0.05, B)
        tr = lambda a: np.ascontiguousarray(a.transpose(0, 2, 1)[pick])
        return x0, tr(g["x_ref"]), tr(g["x_lin"]), tr(g["u_lin"])
    x0, xr, xl, ul = wl.perturbed_batch(model, track, B, seed)
    c = lambda a: np.ascontiguousarray(a[:, :N])
    return x0, c(xr), c(xl), c(ul)


def time_dev(model, track, B, N, steps=3, warm=2):
    mid = fm.KINEMATIC if model == "kinematic" else fm.DYNAMIC
    NX, NU, NS = (5, 2, 1) if model == "kinematic" else (7, 2, 4)
    x0, xr, xl, ul = horizon_batch(model, track, B, N, 100 + rank)
    d = [torch.from_numpy(a).to(dev) for a in (x0, xr, xl, ul)]
    tid_t = torch.full((B,), tnames.index(track), dtype=torch.int32, device=dev)
    pid_t = torch.full((B,), 1 if model == "dynamic" else 0, dtype=torch.int32, device=dev)
    o = dict(u_opt=torch.empty((B, NU * N), dtype=torch.float64, device=dev), x_opt=torch.empty((B, NX * N), dtype=torch.float64, device=dev),
             exitflag=torch.empty(B, dtype=torch.int32, device=dev), fval=torch.empty(B, dtype=torch.float64, device=dev),
             slack_opt=torch.empty((B, NS), dtype=torch.float64, device=dev), iters=torch.empty(B, dtype=torch.int32, device=dev))
    ptrs = dict(x0=d[0].data_ptr(), x_ref=d[1].data_ptr(), x_lin=d[2].data_ptr(), u_lin=d[3].data_ptr(), track_id=tid_t.data_ptr(),
                param_id=pid_t.data_ptr(), **{k: v.data_ptr() for k, v in o.items()})
    for _ in range(warm):
        mpc.ltvmpc_dev(mid, B, N, DT, ptrs, stream=mpc.stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        mpc.ltvmpc_dev(mid, B, N, DT, ptrs, stream=mpc.stream)
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # end to end through the host-buffer API (pageable numpy arrays in, numpy arrays out)
    step = mpc.ltvmpc_kinetmatic_curvilinear if model == "kinematic" else mpc.ltvmpc_dynamic_curvilinear
    tid = np.full(B, tnames.index(track), np.int32); pid = np.full(B, 1 if model == "dynamic" else 0, np.int32)
    step(x0, xr, DT, xl, ul, track_id=tid, param_id=pid)
    t0 = time.perf_counter(); r = step(x0, xr, DT, xl, ul, track_id=tid, param_id=pid); t1 = time.perf_counter()
    return dict(model=model, track=track, horizon=N, batch_per_gpu=B, ms_per_step=ms, qp_per_s=B / ms * 1e3,
                e2e_host_api_qp_per_s=B / (t1 - t0), exitflag_nonzero=int((o["exitflag"] != 0).sum().item()),
                exitflag_infeasible=int((o["exitflag"] == -2).sum().item()), exitflag_max_iter=int((o["exitflag"] == 1).sum().item()),
                iters_mean=float(o["iters"].double().mean().item()))


out = []
# configs[2]
r = time_dev("dynamic", "fss2019", 4096 if quick else 32768, 40)
out.append({"config": "configs[2] dynamic model, fss2019, 262,144 problems over 8 GPUs = 32,768 per GPU", **r})
# configs[3]
B = 2048 if quick else 16384
x0, xr, xl, ul = wl.perturbed_batch("kinematic", "fso2020", B, 300 + rank)
tid = np.full(B, tnames.index("fso2020"), np.int32)
n_sqp = 3
mpc.ltvmpc_sqp(fm.KINEMATIC, x0, xr, DT, xl, ul, n_sqp, track_id=tid)
l0 = mpc.launch_count
t0 = time.perf_counter(); rs = mpc.ltvmpc_sqp(fm.KINEMATIC, x0, xr, DT, xl, ul, n_sqp, track_id=tid); t1 = time.perf_counter()
out.append({"config": "configs[3] SQP: repeated relinearise+QP, batch 16,384, fso2020", "n_sqp": n_sqp, "batch_per_gpu": B,
            "e2e_host_api_qp_per_s": B * n_sqp / (t1 - t0), "e2e_host_api_ms": (t1 - t0) * 1e3, "launches": mpc.launch_count - l0,
            "exitflag_nonzero": int((rs.exitflag != 0).sum())})
# configs[0] scaled out: main.m's closed loop (projection, reference, fused MPC step, actuator PIDs, plant) for B vehicles
Bv, n_sim = (1024, 20) if quick else (16384, 100)
rng = np.random.default_rng(7 + rank)
plant0 = np.zeros((Bv, 7))
plant0[:, 1] = rng.uniform(-0.3, 0.3, Bv)           # lateral offset from the centre line at the start
plant0[:, 2] = rng.uniform(-0.05, 0.05, Bv)
mpc.closed_loop(fm.KINEMATIC, plant0[:64], 5, history=False)
l0 = mpc.launch_count
t0 = time.perf_counter(); cl = mpc.closed_loop(fm.KINEMATIC, plant0, n_sim, history=False); t1 = time.perf_counter()
out.append({"config": "configs[0] as a batch: main.m closed loop on fsg2019 for B vehicles (kinematic LTV-MPC in the loop)",
            "vehicles": Bv, "sim_steps": n_sim, "wall_ms": (t1 - t0) * 1e3, "mpc_steps_per_s": Bv * n_sim / (t1 - t0),
            "launches": mpc.launch_count - l0, "steps_done_min": int(cl["steps"].min())})
# configs[4]
sweep = []
for N in (20, 40, 80):
    for track in (tnames if N != 80 else ["fsg2019"]):
        B = (1024 if N == 80 else 4096) if quick else (8192 if N == 80 else 65536)
        sweep.append(time_dev("kinematic", track, B, N))
tot = sum(s["batch_per_gpu"] for s in sweep); tms = sum(s["ms_per_step"] for s in sweep)
out.append({"config": "configs[4] sweep: horizons 20/40/80 x tracks (N = 20: first 20 steps of the lap problems; N = 80: fsg2019 lap)",
            "cases": sweep, "total_qp": tot, "total_ms": tms, "aggregate_qp_per_s": tot / tms * 1e3})
if rank == 0:
    for o_ in out:
        o_["n_gpus_in_job"] = world
        print(json.dumps(o_))
