#!/usr/bin/env python
"""bench.py -- LTV-MPC QP solves/sec (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

A "step" is one pass of the fused LTV-MPC step (linearise -> condense -> QP solve) over one
batch of B synthetic problems: BASELINE.json configs[1], "batched LTV-MPC kinematic model,
65,536 perturbed initial states on fsg2019, default horizon (40)".  At N > 1 every rank
solves its own B problems (weak scaling, no data-path collective).

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the meaning of every key.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "LTV-MPC QP solves/sec"
UNIT = "QP/s"
WORKLOAD = "kinematic LTV-MPC, fsg2019, N_steps=40, perturbed initial states"
DT = 0.05
N_STEPS = 40


def algorithmic_flops(adds, drops, refreshes, n_qp, nV=81, nU=80, N=40, q_mean=None):
    """FP64 flops the solve path NEEDS (FMA = 2), from the active-set event counts the kernel
    reports (DESIGN.md 'Algorithmic work').  Per QP:
       setup   = B_bar chains + Gramian/Riccati recursions + adjoint rows of J + tile fill
                 + packed H + g + row norms + initial point      (no dense factorisation)
       add     = M'n (nV^2) + J2 y (nV (nV-q)) + rank-1 update (nV^2) + constraint evaluation
       drop    = H k (nV^2) + K1'w (nV q) + K1 update (nV q) + re-projection M'n (nV^2)
       refresh = H x (nV^2) + M'grad (nV^2) + J2 y (nV (nV-q))
    """
    NX, NU, NREAL = 5, 2, 3
    if q_mean is None:
        q_mean = max(1.0, 0.5 * (adds - drops) / max(n_qp, 1))
    cons_eval = nU * (N + 1) / 2 * 2 + 2 * N * (N + 1) / 2          # packed n-rows + 2 prefix sums
    per_stage = (2 * NX * NX * NREAL + 2 * NU * NX * NX               # W A, P A, B'P, B'W
                 + NU * NX * NREAL + 3 * NX + 2 * NU * NX             # S, Lambda, K
                 + 2 * NX * NX * NREAL + NX * NX * NU)                # W', P'
    setup = (nU * (N + 1) / 2 * NREAL * NX                            # B_bar chains
             + N * per_stage                                          # Gramian + Riccati recursions
             + NU * (N * (N - 1) / 2) * (NU * NX + NX * (NREAL + NU)) # adjoint rows of J
             + 2 * nU * (nU + 2) / 4 * 2                              # tile fill (x Lambda^(-T/2))
             + nU * (nU + 1) / 2 * (NREAL + 1)                        # packed H, one dot product per entry
             + nU * (N + 1) / 2 * (NREAL + 1) + nU * (N + 1) / 2 * 2  # g; per-step Gram / column sums
             + 2 * nV * nV)                                           # x0 = -J J' g
    add = 2 * nV * nV + nV * (nV - q_mean) + cons_eval
    drop = 2 * nV * nV + 2 * nV * q_mean
    refresh = 2 * nV * nV + nV * (nV - q_mean)
    fma = n_qp * setup + adds * add + drops * drop + refreshes * refresh
    return 2.0 * fma


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    self.rows.append([c.strip() for c in out.stdout.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


# --------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path (oracle restatement: MATLAB and the
    qpOASES MEX binaries cannot run here), all host cores, bounded sample per step."""
    if rank != 0:
        return
    from fsae_mpc_b200 import workload as wl
    import cpu_baseline
    sample = args.ref_sample
    x0, xr, xl, ul = wl.perturbed_batch("kinematic", "fsg2019", sample, seed=1234)
    tracks = wl.load_tracks()
    base = cpu_baseline.Baseline(tracks["fsg2019"])
    for _ in range(args.warmup):
        base.run(x0[:max(1, sample // 8)], xr, xl, ul, DT)
    t0 = time.perf_counter()
    n_done = 0
    for _ in range(args.steps):
        n_done += base.run(x0, xr, xl, ul, DT)
    el = time.perf_counter() - t0
    v = n_done / el
    line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "batch_per_gpu": args.batch,
                       "sample": f"{sample} problems per step (bounded sample of the {args.batch}-problem batch)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": base.cores, "kind": base.kind,
                             "sample": f"{sample} problems x {args.steps} steps"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------- our arm
def run_ours(args, rank, local_rank, world):
    import torch
    import fsae_mpc_b200 as fm
    from fsae_mpc_b200 import workload as wl

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- fsae_mpc_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("FSAE_NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    mpc = fm.FsaeMpc(local_rank)
    tracks = wl.load_tracks()
    for tid, name in enumerate(tracks):
        t = tracks[name]
        mpc.set_track(tid, t[0], t[1], t[2])
    x0, xr, xl, ul = wl.perturbed_batch("kinematic", "fsg2019", B, seed=1000 + rank)
    NX, NU, NS, N = 5, 2, 1, N_STEPS
    nU, nV, nC = NU * N, NU * N + NS, 6 * N

    # pinned host buffers (inputs and results) for the end-to-end leg
    def pin(a):
        t = torch.from_numpy(a).pin_memory()
        return t
    h_in = [pin(a) for a in (x0, xr, xl, ul)]
    h_out = dict(u_opt=torch.empty((B, nU), dtype=torch.float64).pin_memory(),
                 x_opt=torch.empty((B, NX * N), dtype=torch.float64).pin_memory(),
                 exitflag=torch.empty(B, dtype=torch.int32).pin_memory(),
                 fval=torch.empty(B, dtype=torch.float64).pin_memory(),
                 slack=torch.empty((B, NS), dtype=torch.float64).pin_memory())
    # device-resident copies for the kernel-only leg
    d_in = [t.to(dev) for t in h_in]
    d_out = dict(u_opt=torch.empty((B, nU), dtype=torch.float64, device=dev),
                 x_opt=torch.empty((B, NX * N), dtype=torch.float64, device=dev),
                 exitflag=torch.empty(B, dtype=torch.int32, device=dev),
                 fval=torch.empty(B, dtype=torch.float64, device=dev),
                 slack=torch.empty((B, NS), dtype=torch.float64, device=dev),
                 iters=torch.empty(B, dtype=torch.int32, device=dev))
    ptrs = dict(x0=d_in[0].data_ptr(), x_ref=d_in[1].data_ptr(), x_lin=d_in[2].data_ptr(), u_lin=d_in[3].data_ptr(),
                u_opt=d_out["u_opt"].data_ptr(), x_opt=d_out["x_opt"].data_ptr(),
                exitflag=d_out["exitflag"].data_ptr(), fval=d_out["fval"].data_ptr(),
                slack_opt=d_out["slack"].data_ptr(), iters=d_out["iters"].data_ptr())
    # all timed work runs on the context's own (non-default) stream; torch events are recorded
    # on that same stream through an ExternalStream handle
    stream = torch.cuda.ExternalStream(mpc.stream, device=dev)

    def step_dev():
        mpc.ltvmpc_dev(fm.KINEMATIC, B, N, DT, ptrs, stream=mpc.stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    fp64_peak = mpc.probe_fp64_tflops()

    # ---------------- kernel-only leg: inputs resident in HBM ----------------
    for _ in range(args.warmup):
        step_dev()
    barrier()
    mpc.counters(reset=True)
    l0 = mpc.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with ClockSampler(local_rank) as clk:
        barrier()
        ev[0].record(stream)
        for k in range(args.steps):
            step_dev()
            ev[k + 1].record(stream)
        barrier()
    launches = mpc.launch_count - l0
    total_ms = ev[0].elapsed_time(ev[-1])
    kern_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    adds, drops, refreshes = mpc.counters()
    n_bad = int((d_out["exitflag"] != 0).sum().item())
    iters_mean = float(d_out["iters"].double().mean().item())

    # ---------------- end-to-end leg: host buffers through the C-ABI host call ----------------
    ext = stream
    lib = mpc._lib
    import ctypes as C
    dp = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_double))
    ip = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_int32))

    def step_e2e():
        rc = lib.fsae_ltvmpc_host(mpc._ctx, fm.KINEMATIC, B, N, DT, None, None,
                                  dp(h_in[0]), dp(h_in[1]), dp(h_in[2]), dp(h_in[3]),
                                  dp(h_out["u_opt"]), dp(h_out["x_opt"]), ip(h_out["exitflag"]),
                                  dp(h_out["fval"]), dp(h_out["slack"]), None, None, None)
        if rc != 0:
            raise RuntimeError(mpc._lib.fsae_last_error(mpc._ctx))
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    ee = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ee[0].record(ext)
    for k in range(args.steps):
        step_e2e()
        ee[k + 1].record(ext)
    barrier()
    e2e_ms = ee[0].elapsed_time(ee[-1])
    e2e_step_ms = [ee[k].elapsed_time(ee[k + 1]) for k in range(args.steps)]

    # ---------------- single-problem step latency (the control-loop view of the metric) -------
    lat_us = []
    if rank == 0:
        one = dict(ptrs)
        for _ in range(20):
            mpc.ltvmpc_dev(fm.KINEMATIC, 1, N, DT, one, stream=mpc.stream)
        torch.cuda.synchronize(dev)
        for _ in range(200):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(ext)
            mpc.ltvmpc_dev(fm.KINEMATIC, 1, N, DT, one, stream=mpc.stream)
            a1.record(ext)
            a1.synchronize()
            lat_us.append(a0.elapsed_time(a1) * 1e3)
    h2d = sum(t.numel() * t.element_size() for t in h_in)
    d2h = sum(t.numel() * t.element_size() for t in h_out.values())

    # ---------------- max over ranks ----------------
    from fsae_mpc_b200.sharding import reduce_metrics
    (total_ms, e2e_ms), sums_l = reduce_metrics([total_ms, e2e_ms],
                                                [adds, drops, refreshes, n_bad], dist, device=dev)
    sums = torch.tensor(sums_l, dtype=torch.float64)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    n_qp = world * B * args.steps
    value = n_qp / (total_ms * 1e-3)
    e2e_value = n_qp / (e2e_ms * 1e-3)
    peaks, peak_src = measured_peaks()
    # roofline of the dominant (only) kernel: FP64 FMA pipe; per-launch = per step
    flops_launch = algorithmic_flops(adds, drops, refreshes, B * args.steps) / args.steps   # rank 0's own counts
    kms = float(np.mean(kern_ms))
    ach_tf = flops_launch / (kms * 1e-3) / 1e12
    alg_bytes = B * 8 * (NX + 2 * NX * N + NU * N + nU + NX * N + NS + 2) + B * 4
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
    if os.path.exists(tp) and B == 65536:
        with open(tp) as fh:
            tj = json.load(fh)
        traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        traffic_src = tj["source"]
    roof = {"bound": "fp64_fma", "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": ach_tf / fp64_peak if fp64_peak else None, "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": "measured in this run (fsae_probe_fp64_tflops, DFMA streams on all SMs)",
            "kernel": "ltvmpc_fused_kernel", "kernel_ms": kms,
            "algorithmic_flops_per_launch": flops_launch,
            "hbm": {"achieved": alg_bytes / (kms * 1e-3) / 1e9, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                    "frac": alg_bytes / (kms * 1e-3) / 1e9 / peaks.get("hbm_gbs", 1), "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes}}

    # CPU baseline on a bounded sample, rank 0, N=1 only
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import cpu_baseline
        base = cpu_baseline.Baseline(tracks["fsg2019"])
        s = args.ref_sample
        t0 = time.perf_counter()
        done = base.run(x0[:s], xr[:s], xl[:s], ul[:s], DT)
        el = time.perf_counter() - t0
        cpu = {"value": done / el, "unit": UNIT, "cores": base.cores, "kind": base.kind,
               "sample": f"first {s} problems of the batch, {el:.1f} s"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": world * B, "horizon": N,
                       "parallelism": f"dp{world} (independent problems, no collective on the solve path)",
                       "l2": f"inputs {h2d / 1e6:.0f} MB per step > 126 MB L2; no flush needed"},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps, "api": "fsae_ltvmpc_host (C-ABI, pinned host buffers)"},
            "latency": {"step_ms_p50": float(np.percentile(kern_ms, 50)), "step_ms_p99": float(np.percentile(kern_ms, 99)),
                        "e2e_step_ms_p50": float(np.percentile(e2e_step_ms, 50)),
                        "e2e_step_ms_p99": float(np.percentile(e2e_step_ms, 99)),
                        "single_problem_us_p50": float(np.percentile(lat_us, 50)) if lat_us else None,
                        "single_problem_us_p99": float(np.percentile(lat_us, 99)) if lat_us else None,
                        "note": "step = one fused pass over batch_per_gpu problems (rank 0); single_problem = one "
                                "MPC step of one vehicle, device-resident, 200 samples"},
            "gpu_launches": int(launches),
            "roofline": roof,
            "cpu_baseline": cpu,
            "solver": {"exitflag_nonzero": int(sums[3].item()), "iters_mean": iters_mean,
                       "adds_per_qp": sums[0].item() / n_qp, "drops_per_qp": sums[1].item() / n_qp,
                       "refreshes_per_qp": sums[2].item() / n_qp}}
    sys.stdout.flush()
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=65536, help="problems per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-sample", type=int, default=None, help="problems per reference/CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.ref_sample is None:
        args.ref_sample = 16384 if args.impl == "reference" else 32768
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
