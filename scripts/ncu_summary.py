"""Summarise an .ncu-rep (read here, no GPU needed): key raw metrics + hottest source lines.
    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [out.md]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "local_load", "local_store", "smsp__inst_executed_op_local",
        ]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
print(f"# ncu summary of {rep}\n", file=out)
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print(f"## kernel: {d.get('Kernel Name','?')[:100]}  (launch id {d.get('ID','?')})\n", file=out)
    print("| metric | unit | value |\n|---|---|---|", file=out)
    for h, u, v in zip(hdr, units, vals):
        if any(k == h or (k in h and 'stalled' not in h and k.startswith(('local','smsp__inst_executed_op_local'))) for k in KEYS):
            print(f"| {h} | {u} | {v} |", file=out)
    print("\nstall reasons (warps per issue-active cycle):\n", file=out)
    st = [(float(v), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
          for h, v in zip(hdr, vals) if h.startswith("smsp__average_warps_issue_stalled_") and v]
    for v, h in sorted(st, reverse=True)[:8]:
        print(f"- {h}: {v:.2f}", file=out)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur = None
lines = []
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] not in ("", "Line No") and r[2] == "-":
        try:
            lines.append((int(r[6]), int(r[7]), cur, r[0], r[1].strip()[:95]))
        except ValueError:
            pass
tot = sum(l[0] for l in lines) or 1
toti = sum(l[1] for l in lines) or 1
print(f"\n## hottest source lines (of {tot} stall samples, {toti} warp instructions)\n", file=out)
print("| samples % | instr % | file:line | source |\n|---|---|---|---|", file=out)
for s, i, f, ln, txt in sorted(lines, reverse=True)[:40]:
    print(f"| {100*s/tot:.1f} | {100*i/toti:.1f} | {f}:{ln} | `{txt}` |", file=out)
