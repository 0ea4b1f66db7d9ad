#!/usr/bin/env python
"""Turn gpurun_out/<round>_* ncu outputs into the small text files committed under profiles/.
usage: python tools/summarise_profiles.py r01"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
OUT = os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)
G = os.path.join(ROOT, "gpurun_out")

# 1. launch list: kernel name, count, total and share of device time
path = os.path.join(G, f"{R}_launches.csv")
if os.path.exists(path):
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}.get(unit, 1e-6)
            rows.append((r["Kernel Name"], v * scale))
    tot = sum(t for _, t in rows) or 1.0
    agg = {}
    for k, t in rows:
        a = agg.setdefault(k.split("(")[0][:90], [0, 0.0])
        a[0] += 1
        a[1] += t
    with open(os.path.join(OUT, f"{R}_launch_list.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none : python bench.py --steps 2 --warmup 3 --no-extra\n")
        f.write(f"# {len(rows)} launches, {tot:.3f} ms of device time (cold-cache, serialised: compare shares)\n")
        f.write(f"{'kernel':90s} {'launches':>8s} {'ms':>10s} {'share':>7s}\n")
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{k:90s} {n:8d} {t:10.3f} {100 * t / tot:6.1f}%\n")
    print("wrote launch list:", len(rows), "launches")

# 2. full capture of the headline kernel
rep = os.path.join(G, f"{R}_lean_full.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    h, u, v = r[0], r[1], r[2]
    keep = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
            "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
    vals = {}
    with open(os.path.join(OUT, f"{R}_lean_kernel_ncu.txt"), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on -k regex:lean2_kernel : CQ_BENCH_BYTES=2e9 python bench.py --steps 2 --warmup 3 --no-extra\n")
        for i, k in enumerate(h):
            if k in keep or (k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")):
                f.write(f"{k} = {v[i]} {u[i]}\n")
                vals[k] = v[i]
    try:
        rd = float(vals["dram__bytes_read.sum"].replace(",", ""))
        wr = float(vals["dram__bytes_write.sum"].replace(",", ""))
        bj = json.load(open(os.path.join(G, f"{R}_bench_plain.json")))
        # the capture ran on a 2e9-byte input; units as ncu printed them
        print("dram read/write as printed:", vals["dram__bytes_read.sum"], vals["dram__bytes_write.sum"])
    except Exception as ex:
        print("traffic:", ex)
    lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "25"], capture_output=True, text=True).stdout
    with open(os.path.join(OUT, f"{R}_lean_kernel_hot_lines.txt"), "w") as f:
        f.write(lines)
    print("wrote lean kernel summary")
