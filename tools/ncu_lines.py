#!/usr/bin/env python
"""Top source lines of a kernel in an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [N]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr = None, None
lines = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_inst = hdr.index("Instructions Executed")
        i_samp = hdr.index("# Samples")
        continue
    if r[0] == "Function Name" or hdr is None:
        continue
    if r[0] == "":
        continue  # SASS rows
    try:
        lines.append((cur_file, int(r[0]), r[1].strip(), int(r[i_inst]), int(r[i_samp])))
    except Exception:
        pass
tot_i = sum(x[3] for x in lines) or 1
tot_s = sum(x[4] for x in lines) or 1
print(f"total warp-instructions {tot_i}, samples {tot_s}")
print("--- by instructions executed")
for f, ln, src, n, sm in sorted(lines, key=lambda x: -x[3])[:top]:
    print(f"{100 * n / tot_i:5.1f}% inst {100 * sm / tot_s:5.1f}% stall  {f}:{ln}: {src[:110]}")
print("--- by stall samples")
for f, ln, src, n, sm in sorted(lines, key=lambda x: -x[4])[:top]:
    print(f"{100 * sm / tot_s:5.1f}% stall {100 * n / tot_i:5.1f}% inst  {f}:{ln}: {src[:110]}")
