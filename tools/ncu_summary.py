#!/usr/bin/env python
"""Key numbers of the first kernel in an .ncu-rep + per-source-line instruction counts.
usage: python tools/ncu_summary.py report.ncu-rep [rows_per_launch]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
h, v = r[0], r[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for i, k in enumerate(h):
    if k in want or k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
        print(f"{k} = {v[i]} {r[1][i]}")
