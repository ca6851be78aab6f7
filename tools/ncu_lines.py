#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics + dynamic instruction count / stall samples per CUDA source line."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print(f"{k:90s} {rows[1][i]:>8s} " + " ".join(r[i] for r in rows[2:]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = None; fpath = None; agg = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fpath = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] != "":
        agg.append((fpath, r))
if hdr:
    iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples")
    num = lambda x: int(x) if x.isdigit() else 0
    tot = sum(num(r[iI]) for f, r in agg); totS = max(1, sum(num(r[iS]) for f, r in agg))
    print("total inst", tot, "samples", totS)
    for f, r in sorted(agg, key=lambda fr: -num(fr[1][iI]))[:top]:
        print(f"{num(r[iI])*100/tot:5.2f}% inst {num(r[iS])*100/totS:5.2f}% smp {f}:{r[0]:>4s} {r[1].strip()[:100]}")
