#!/usr/bin/env python
"""Static SASS instruction count per source line of a cubin compiled with -lineinfo (nvdisasm -g -c):
   python tools/sass_lines.py kernel.cubin [file-substring] [first-line last-line]"""
import re
import subprocess
import sys
from collections import Counter

cubin = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4]) if len(sys.argv) > 4 else 10 ** 9
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
cur = None
per = Counter()
ops = Counter()
total = 0
for line in dis.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        if "inlined at" in line and cur is not None:
            continue
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
    if m and cur:
        total += 1
        per[cur] += 1
        if want in cur[0] and lo <= cur[1] <= hi:
            ops[m.group(1).split(".")[0]] += 1
print("total", total)
sel = 0
for (f, l), n in sorted(per.items()):
    if want in f and lo <= l <= hi:
        print(f"{f}:{l}\t{n}")
        sel += n
print("selected", sel)
print(ops.most_common(30))
