#!/usr/bin/env python
"""Turn gpurun_out/<round>_* outputs of tools/collect_profiles.sh into the small text files committed under profiles/.
usage: python tools/summarise_profiles.py r02"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
OUT = os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)
G = os.path.join(ROOT, "gpurun_out")

# 0. the bench line of the same run
for src, dst in ((f"{R}_bench_plain.json", f"{R}_bench_line.json"),):
    path = os.path.join(G, src)
    if os.path.exists(path):
        line = [ln for ln in open(path) if ln.startswith("{")][-1]
        json.dump(json.loads(line), open(os.path.join(OUT, dst), "w"), indent=1)
        print("wrote", dst)

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
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 600 : python bench.py --steps 2 --warmup 3 --no-extra\n")
        f.write(f"# {len(rows)} launches, {tot:.3f} ms of device time (cold-cache, serialised: compare shares). The first launches are the\n")
        f.write("# synthetic-data generator; then the three timed legs (lean2k, lean2<ONELEAF>, leanhc + expand + device-side finish) and the end-to-end legs (lean2k again)\n")
        f.write(f"{'kernel':90s} {'launches':>8s} {'ms':>10s} {'share':>7s}\n")
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{k:90s} {n:8d} {t:10.3f} {100 * t / tot:6.1f}%\n")
    print("wrote launch list:", len(rows), "launches")

# 2. full captures
KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_atom.sum",
        "lts__t_sectors_srcunit_tex_op_red.sum", "sm__cycles_elapsed.avg.per_second"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
traffic = {}
for kern, plan in (("lean2k", "group_name"), ("lean2", "count_age_gt_40"), ("leanhc", "group_high_card")):
    rep = os.path.join(G, f"{R}_{kern}.ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    h, u, v = r[0], r[1], r[2]
    try:  # DRAM bytes of the captured launch per byte of its table (tools/run_plan.py prints the table size: 2e9 / 29.89 rows)
        rd = float(v[h.index("dram__bytes_read.sum")].replace(",", "")) * UNIT[u[h.index("dram__bytes_read.sum")]]
        wr = float(v[h.index("dram__bytes_write.sum")].replace(",", "")) * UNIT[u[h.index("dram__bytes_write.sum")]]
        plain = open(os.path.join(G, f"{R}_{kern}_plain.log")).read()
        gbs = float(plain.strip().splitlines()[-1].split("GB/s")[1].split()[0])
        kms = float(plain.strip().splitlines()[-1].split("kernel_ms")[1].split()[0])
        table_bytes = gbs * 1e9 * kms / 1e3
        traffic[kern] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "table_bytes": table_bytes,
                         "dram_bytes_per_csv_byte": (rd + wr) / table_bytes, "capture": f"profiles/{R}_{kern}_ncu.txt"}
    except Exception as ex:
        print("traffic", kern, ex)
    with open(os.path.join(OUT, f"{R}_{kern}_ncu.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on -k regex:{kern} : python tools/run_plan.py {plan} 2e9 3\n")
        f.write("# (one launch on a 2 GB table, kernel compiled for the query by cqg_jit where that applies; tools/collect_profiles.sh)\n")
        for i, k in enumerate(h):
            if k in KEEP or (k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")):
                f.write(f"{k} = {v[i]} {u[i]}\n")
        f.write("\n")
        f.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "25"], capture_output=True, text=True).stdout)
    print("wrote", f"{R}_{kern}_ncu.txt")
if traffic:
    json.dump(traffic, open(os.path.join(OUT, f"{R}_traffic.json"), "w"), indent=1)
    print("wrote traffic ratios", {k: round(v["dram_bytes_per_csv_byte"], 4) for k, v in traffic.items()})
