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


def algorithmic_flops(adds, drops, refreshes, n_qp, nV=81, nU=80, N=40, q_mean=None, NX=5, NREAL=3, NCR=1, NROWS=160):
    """FP64 flops the solve path NEEDS (FMA = 2), from the active-set event counts the kernel
    reports (DESIGN.md 'Algorithmic work').  Per QP:
       setup   = B_bar chains + Gramian/Riccati recursions + adjoint rows of J + tile fill
                 + packed H + g + row norms + initial point      (no dense factorisation)
       add     = M'n (nV^2) + J2 y (nV (nV-q)) + rank-1 update (nV^2) + constraint evaluation
       drop    = H k (nV^2) + K1'w (nV q) + K1 update (nV q) + re-projection M'n (nV^2)
       refresh = H x (nV^2) + M'grad (nV^2) + J2 y (nV (nV-q))
    """
    NU = 2
    if q_mean is None:
        q_mean = max(1.0, 0.5 * (adds - drops) / max(n_qp, 1))
    cons_eval = NCR * nU * (N + 1) / 2 + 2 * N * (N + 1) / 2 + 4 * NROWS   # packed constraint rows + prefix sums + row forms
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


def make_config(world, B):
    """The workload description BOTH arms print (the driver compares the two lines' `config`)."""
    return {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": world * B, "horizon": N_STEPS,
            "seed": "1000 + rank (fsae_mpc_b200.workload.perturbed_batch)",
            "parallelism": f"dp{world} (independent problems, no collective on the solve path)",
            "l2": "inputs 254 MB per step > 126 MB L2; no flush needed"}


# --------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path (oracle restatement: MATLAB and the qpOASES MEX
    binaries cannot run here), all host cores, on THE SAME problems as our arm (rank 0's batch: same generator,
    same seed).  A step solves the first `sample` problems of that batch; `sample` is the whole batch when the
    host is fast enough for the run to end within a few minutes, else the largest multiple of 4096 that is."""
    if rank != 0:
        return
    from fsae_mpc_b200 import workload as wl
    import cpu_baseline
    B = args.batch
    x0, xr, xl, ul = wl.perturbed_batch("kinematic", "fsg2019", B, seed=1000)
    tracks = wl.load_tracks()
    base = cpu_baseline.Baseline(tracks["fsg2019"])
    n_w = max(1, min(B, 2048))
    t0 = time.perf_counter()
    for _ in range(max(1, args.warmup)):
        base.run(x0[:n_w], xr[:n_w], xl[:n_w], ul[:n_w], DT)
    rate = max(1, args.warmup) * n_w / (time.perf_counter() - t0)          # problems / s on this host
    sample = args.ref_sample
    if sample is None:
        budget_s = 150.0
        sample = int(min(B, max(4096, (rate * budget_s / max(1, args.steps)) // 4096 * 4096)))
    sample = min(sample, B)
    t0 = time.perf_counter()
    n_done = 0
    for _ in range(args.steps):
        n_done += base.run(x0[:sample], xr[:sample], xl[:sample], ul[:sample], DT)
    el = time.perf_counter() - t0
    v = n_done / el
    what = "the whole batch" if sample == B else f"the first {sample} problems of the {B}-problem batch"
    line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": make_config(world, B),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": base.cores, "kind": base.kind,
                             "sample": f"{what} per step x {args.steps} steps ({el:.1f} s)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------- our arm
class Leg:
    """One workload resident on one GPU: host (pinned + pageable) and device buffers, the three ways of running it."""

    def __init__(self, torch, fm, mpc, dev, model, x0, xr, xl, ul, track_id=0, param_id=0):
        import ctypes as C
        self.torch, self.fm, self.mpc, self.dev = torch, fm, mpc, dev
        self.mid = fm.KINEMATIC if model == "kinematic" else fm.DYNAMIC
        self.NX, self.NU, self.NS = (5, 2, 1) if model == "kinematic" else (7, 2, 4)
        B, N = x0.shape[0], xr.shape[1]
        self.B, self.N = B, N
        NX, NU, NS = self.NX, self.NU, self.NS
        self.np_in = [x0, xr, xl, ul]                                      # pageable (plain numpy)
        self.ids = (np.full(B, track_id, np.int32), np.full(B, param_id, np.int32))
        self.h_in = [torch.from_numpy(a).pin_memory() for a in self.np_in]
        shapes = dict(u_opt=((B, NU * N), torch.float64), x_opt=((B, NX * N), torch.float64), exitflag=((B,), torch.int32),
                      fval=((B,), torch.float64), slack=((B, NS), torch.float64))
        self.h_out = {k: torch.empty(sh, dtype=dt).pin_memory() for k, (sh, dt) in shapes.items()}
        self.p_out = {k: np.empty(sh, dtype=np.float64 if dt == torch.float64 else np.int32) for k, (sh, dt) in shapes.items()}
        self.d_in = [t.to(dev) for t in self.h_in]
        self.d_ids = [torch.from_numpy(a).to(dev) for a in self.ids]
        self.d_out = {k: torch.empty(sh, dtype=dt, device=dev) for k, (sh, dt) in shapes.items()}
        self.d_out["iters"] = torch.empty(B, dtype=torch.int32, device=dev)
        self.ptrs = dict(x0=self.d_in[0].data_ptr(), x_ref=self.d_in[1].data_ptr(), x_lin=self.d_in[2].data_ptr(),
                         u_lin=self.d_in[3].data_ptr(), track_id=self.d_ids[0].data_ptr(), param_id=self.d_ids[1].data_ptr(),
                         u_opt=self.d_out["u_opt"].data_ptr(), x_opt=self.d_out["x_opt"].data_ptr(),
                         exitflag=self.d_out["exitflag"].data_ptr(), fval=self.d_out["fval"].data_ptr(),
                         slack_opt=self.d_out["slack"].data_ptr(), iters=self.d_out["iters"].data_ptr())
        self.h2d = sum(t.numel() * t.element_size() for t in self.h_in)
        self.d2h = sum(t.numel() * t.element_size() for t in self.h_out.values())
        self._C = C

    def step_dev(self):
        self.mpc.ltvmpc_dev(self.mid, self.B, self.N, DT, self.ptrs, stream=self.mpc.stream)

    def _host(self, ins, outs, addr):
        C = self._C
        dp = lambda t: C.cast(addr(t), C.POINTER(C.c_double))
        ip = lambda t: C.cast(addr(t), C.POINTER(C.c_int32))
        ipn = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        rc = self.mpc._lib.fsae_ltvmpc_host(self.mpc._ctx, self.mid, self.B, self.N, DT, ipn(self.ids[0]), ipn(self.ids[1]),
                                            dp(ins[0]), dp(ins[1]), dp(ins[2]), dp(ins[3]),
                                            dp(outs["u_opt"]), dp(outs["x_opt"]), ip(outs["exitflag"]),
                                            dp(outs["fval"]), dp(outs["slack"]), None, None, None)
        if rc != 0:
            raise RuntimeError(self.mpc._lib.fsae_last_error(self.mpc._ctx))

    def step_e2e(self):            # pinned host buffers through the C-ABI host call
        self._host(self.h_in, self.h_out, lambda t: t.data_ptr())

    def step_pageable(self):       # pageable host buffers (what a MEX gateway passes): pinned staging ring inside the library
        self._host(self.np_in, self.p_out, lambda a: a.ctypes.data)

    def timed(self, fn, steps, warmup, stream, barrier):
        torch = self.torch
        for _ in range(warmup):
            fn()
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record(stream)
        for k in range(steps):
            fn()
            ev[k + 1].record(stream)
        barrier()
        return ev[0].elapsed_time(ev[-1]), [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]


def run_ours(args, rank, local_rank, world):
    import torch
    import fsae_mpc_b200 as fm
    from fsae_mpc_b200 import workload as wl
    from fsae_mpc_b200.sharding import reduce_metrics, shard_range

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- fsae_mpc_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")    # never override what the launcher set; NCCL logs go to stderr
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    mpc = fm.FsaeMpc(local_rank)
    tracks = wl.load_tracks()
    for tid, name in enumerate(tracks):
        t = tracks[name]
        mpc.set_track(tid, t[0], t[1], t[2])
    mpc.set_params(1, fm.default_params(fm.DYNAMIC))
    NX, NU, NS, N = 5, 2, 1, N_STEPS
    nU, nV = NU * N, NU * N + NS
    kin = Leg(torch, fm, mpc, dev, "kinematic", *wl.perturbed_batch("kinematic", "fsg2019", B, seed=1000 + rank))
    # all timed work runs on the context's own (non-default) stream; torch events are recorded
    # on that same stream through an ExternalStream handle
    stream = torch.cuda.ExternalStream(mpc.stream, device=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    fp64_peak = mpc.probe_fp64_tflops()

    # ---------------- kernel-only leg: inputs resident in HBM ----------------
    for _ in range(args.warmup):
        kin.step_dev()
    barrier()
    mpc.counters(reset=True)
    l0 = mpc.launch_count
    with ClockSampler(local_rank) as clk:
        total_ms, kern_ms = kin.timed(kin.step_dev, args.steps, 0, stream, barrier)
    launches = mpc.launch_count - l0
    adds, drops, refreshes = mpc.counters()
    n_bad = int((kin.d_out["exitflag"] != 0).sum().item())
    iters_mean = float(kin.d_out["iters"].double().mean().item())

    # ---------------- end-to-end legs: host buffers through the C-ABI host call ----------------
    e2e_ms, e2e_step_ms = kin.timed(kin.step_e2e, args.steps, max(1, args.warmup // 2), stream, barrier)
    pg_ms, _ = kin.timed(kin.step_pageable, args.steps, max(1, args.warmup // 2), stream, barrier)
    staged = mpc.last_host_path == 1
    same = bool(np.array_equal(kin.p_out["u_opt"], kin.h_out["u_opt"].numpy()))

    # ---------------- dynamic model (the reference's default, main.m:26; BASELINE configs[2]) ----------------
    Bd = args.dyn_batch
    dyn_res = None
    if Bd > 0:
        dyn = Leg(torch, fm, mpc, dev, "dynamic", *wl.perturbed_batch("dynamic", "fss2019", Bd, seed=2000 + rank), track_id=1, param_id=1)
        for _ in range(3):
            dyn.step_dev()
        barrier()
        mpc.counters(reset=True)
        d_ms, d_kern = dyn.timed(dyn.step_dev, args.steps, 0, stream, barrier)
        d_adds, d_drops, d_refr = mpc.counters()
        d_e2e_ms, _ = dyn.timed(dyn.step_e2e, args.steps, 1, stream, barrier)
        d_pg_ms, _ = dyn.timed(dyn.step_pageable, args.steps, 1, stream, barrier)
        d_bad = int((dyn.d_out["exitflag"] != 0).sum().item())
        d_iters = float(dyn.d_out["iters"].double().mean().item())
        dyn_res = dict(ms=d_ms, kern=d_kern, e2e=d_e2e_ms, pg=d_pg_ms, adds=d_adds, drops=d_drops, refr=d_refr, bad=d_bad,
                       iters=d_iters, h2d=dyn.h2d, d2h=dyn.d2h)
        del dyn

    # ---------------- strong scaling: ONE 65,536-problem batch split over the ranks ----------------
    strong = None
    if world > 1:
        lo, hi = shard_range(B, rank, world)
        g = wl.perturbed_batch("kinematic", "fsg2019", B, seed=1000)      # the N = 1 batch, identical on every rank
        sl = Leg(torch, fm, mpc, dev, "kinematic", *(np.ascontiguousarray(a[lo:hi]) for a in g))
        s_ms, _ = sl.timed(sl.step_dev, args.steps, 3, stream, barrier)
        s_e2e, _ = sl.timed(sl.step_e2e, args.steps, 1, stream, barrier)
        strong = (s_ms, s_e2e)
        del sl

    # ---------------- single-problem step latency (the control-loop view of the metric) -------
    lat_us = []
    if rank == 0:
        for _ in range(20):
            mpc.ltvmpc_dev(fm.KINEMATIC, 1, N, DT, kin.ptrs, stream=mpc.stream)
        torch.cuda.synchronize(dev)
        for _ in range(200):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            mpc.ltvmpc_dev(fm.KINEMATIC, 1, N, DT, kin.ptrs, stream=mpc.stream)
            a1.record(stream)
            a1.synchronize()
            lat_us.append(a0.elapsed_time(a1) * 1e3)
    h2d, d2h = kin.h2d, kin.d2h

    # ---------------- max over ranks ----------------
    tms = [total_ms, e2e_ms, pg_ms] + ([dyn_res["ms"], dyn_res["e2e"], dyn_res["pg"]] if dyn_res else [0, 0, 0]) + (list(strong) if strong else [0, 0])
    cnt = [adds, drops, refreshes, n_bad] + ([dyn_res["adds"], dyn_res["drops"], dyn_res["refr"], dyn_res["bad"]] if dyn_res else [0, 0, 0, 0])
    tms, sums_l = reduce_metrics(tms, cnt, dist, device=dev)
    total_ms, e2e_ms, pg_ms = tms[:3]
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
    for tp in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
        tp = os.path.join(ROOT, "profiles", tp)
        if os.path.exists(tp) and B == 65536:
            with open(tp) as fh:
                tj = json.load(fh)
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
            traffic_src = tj["source"]
            break
    roof = {"bound": "fp64_fma", "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": ach_tf / fp64_peak if fp64_peak else None, "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": "measured in this run (fsae_probe_fp64_tflops, DFMA streams on all SMs)",
            "kernel": "ltvmpc_fused_v2_kernel<KinModel,40,...>", "kernel_ms": kms,
            "algorithmic_flops_per_launch": flops_launch,
            "hbm": {"achieved": alg_bytes / (kms * 1e-3) / 1e9, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                    "frac": alg_bytes / (kms * 1e-3) / 1e9 / peaks.get("hbm_gbs", 1), "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes}}

    # CPU baseline on a bounded sample, rank 0, N=1 only
    cpu, cpu_dyn = None, None
    if world == 1 and not args.no_cpu_baseline:
        import cpu_baseline
        base = cpu_baseline.Baseline(tracks["fsg2019"])
        s = args.ref_sample or 32768
        x0, xr, xl, ul = kin.np_in
        t0 = time.perf_counter()
        done = base.run(x0[:s], xr[:s], xl[:s], ul[:s], DT)
        el = time.perf_counter() - t0
        cpu = {"value": done / el, "unit": UNIT, "cores": base.cores, "kind": base.kind,
               "sample": f"first {s} problems of the batch, {el:.1f} s"}
        if dyn_res and hasattr(cpu_baseline, "DynamicBaseline"):
            based = cpu_baseline.DynamicBaseline(tracks["fss2019"])
            sd = min(Bd, 4096)
            xd = wl.perturbed_batch("dynamic", "fss2019", Bd, seed=2000)
            t0 = time.perf_counter()
            done = based.run(xd[0][:sd], xd[1][:sd], xd[2][:sd], xd[3][:sd], DT)
            el = time.perf_counter() - t0
            cpu_dyn = {"value": done / el, "unit": UNIT, "cores": based.cores, "kind": based.kind,
                       "sample": f"first {sd} problems of the dynamic batch, {el:.1f} s"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": make_config(world, B),
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps, "api": "fsae_ltvmpc_host (C-ABI, pinned host buffers)"},
            "e2e_pageable": {"value": n_qp / (pg_ms * 1e-3), "unit": UNIT, "ms_per_step": pg_ms / args.steps,
                             "over_device": (n_qp / (pg_ms * 1e-3)) / value, "staging_ring_used": staged,
                             "bit_identical_to_pinned_path": same,
                             "api": "fsae_ltvmpc_host (C-ABI, PAGEABLE host buffers as a MEX gateway passes them; pinned "
                                    "staging ring + copy threads inside the library)"},
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
            "solver": {"exitflag_nonzero": int(sums_l[3]), "iters_mean": iters_mean,
                       "adds_per_qp": sums_l[0] / n_qp, "drops_per_qp": sums_l[1] / n_qp,
                       "refreshes_per_qp": sums_l[2] / n_qp}}
    if dyn_res:
        nd = world * Bd * args.steps
        dk = float(np.mean(dyn_res["kern"]))
        dfl = algorithmic_flops(dyn_res["adds"], dyn_res["drops"], dyn_res["refr"], Bd * args.steps, nV=84, nU=80, N=40,
                                NX=7, NREAL=6, NCR=4, NROWS=680) / args.steps
        line["dynamic"] = {"workload": "dynamic (tyre-force) LTV-MPC, fss2019, N_steps=40, perturbed initial states (BASELINE configs[2]: "
                                       "262,144 problems on 8 GPUs = 32,768 per GPU)",
                           "batch_per_gpu": Bd, "global_batch": world * Bd, "value": nd / (tms[3] * 1e-3), "unit": UNIT,
                           "ms_per_step": tms[3] / args.steps,
                           "e2e": {"value": nd / (tms[4] * 1e-3), "unit": UNIT, "h2d_bytes_per_step": dyn_res["h2d"],
                                   "d2h_bytes_per_step": dyn_res["d2h"]},
                           "e2e_pageable": {"value": nd / (tms[5] * 1e-3), "unit": UNIT, "over_device": tms[3] / tms[5]},
                           "roofline": {"bound": "fp64_fma", "achieved": dfl / (dk * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                                        "frac": dfl / (dk * 1e-3) / 1e12 / fp64_peak if fp64_peak else None,
                                        "kernel": "ltvmpc_fused_v2_kernel<DynModel,40,...>", "kernel_ms": dk,
                                        "algorithmic_flops_per_launch": dfl},
                           "cpu_baseline": cpu_dyn,
                           "solver": {"exitflag_nonzero": int(sums_l[7]), "iters_mean": dyn_res["iters"],
                                      "adds_per_qp": sums_l[4] / nd, "drops_per_qp": sums_l[5] / nd, "refreshes_per_qp": sums_l[6] / nd}}
    if strong:
        ns = B * args.steps
        line["strong_scaling"] = {"global_batch": B, "batch_per_gpu": f"{B // world} (+1 on the first {B % world} ranks)" if B % world else B // world,
                                  "value": ns / (tms[6] * 1e-3), "unit": UNIT, "ms_per_step": tms[6] / args.steps,
                                  "e2e": {"value": ns / (tms[7] * 1e-3), "unit": UNIT},
                                  "note": "the N = 1 batch (seed 1000) split over the ranks with shard_range; max over ranks"}
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
    ap.add_argument("--dyn-batch", type=int, default=32768, help="dynamic-model problems per GPU per step (0: skip that leg)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-sample", type=int, default=None, help="problems per reference/CPU-baseline step (default: adaptive)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
