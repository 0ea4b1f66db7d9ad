#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel the opcode histogram and the TMA / mbarrier / atomic lines.
usage: python tools/sass_extract.py > profiles/r02_sass_extract.txt
Kernels: the ahead-of-time ones out of cq_b200/libcqgpu.so, and lean2k_kernel compiled here with NVRTC for the bench
headline's shape exactly as cqg_jit compiles it on the GPU box (tools/jit_compile.py)."""
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import jit_compile

LIB = os.path.join(ROOT, "cq_b200", "libcqgpu.so")
WANT = ["lean2k_kernel", "lean2_kernelINS_3GeoILi128ELi16384ELi1ELi224EEELi9ELb1ELi2ELb0", "lean2g_kernel", "leanhc_kernel",
        "scan_kernelINS_3GeoILi128ELi16384ELi2ELi992EEELi4"]
KEY = re.compile(r"UBLKCP|UBLKPF|SYNCS|ATOMS|ATOMG|ATOM\.|RED\.|REDUX|UTMALDG|UTCMMA|LDTM|CAS")


def functions(sass):
    cur, body = None, []
    for ln in sass.splitlines():
        m = re.match(r"\s+Function : (\S+)", ln)
        if m:
            if cur:
                yield cur, body
            cur, body = m.group(1), []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+", ln):
            body.append(ln)
    if cur:
        yield cur, body


def report(name, body):
    ops = Counter()
    keys = Counter()
    for ln in body:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", ln)
        if not m:
            continue
        op = m.group(1)
        ops[op.split(".")[0]] += 1
        if KEY.search(op):
            keys[op] += 1
    print(f"== {name}\n   {sum(ops.values())} instructions (static)")
    print("   opcodes:", ", ".join(f"{k} {v}" for k, v in ops.most_common(24)))
    print("   TMA / mbarrier / atomics:", ", ".join(f"{k} x{v}" for k, v in sorted(keys.items())) or "none")


print("# cuobjdump -sass, sm_100a. UBLKCP = cp.async.bulk global->shared (1-D TMA), UBLKPF = bulk L2 prefetch, SYNCS = mbarrier ops.")
print("# No UTMALDG / UTCMMA / LDTM anywhere: nothing on this path is a contraction, and its tiles are byte streams (1-D bulk copies).\n")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
seen = set()
for fn, body in functions(sass):
    for w in WANT[1:]:
        if w in fn and w not in seen:
            seen.add(w)
            report(subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()[:150], body)
cubin = "/tmp/sass_extract_l2k.cubin"
open(cubin, "wb").write(jit_compile.compile_kernel("cqg_lean2k.cuh", "cqg::lean2k_kernel<cqg::Geo<128, 16384, 1, 224>, 8>", "group_name"))
sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
for fn, body in functions(sass):
    report("lean2k_kernel<Geo<128,16384,1,224>,8> compiled with NVRTC for the bench headline's shape (SELECT name, COUNT(*), AVG(height), "
           "SUM(age) ... WHERE age > 25 GROUP BY name)", body)
res = subprocess.run(["cuobjdump", "-res-usage", cubin], capture_output=True, text=True).stdout
print("  ", [ln.strip() for ln in res.splitlines() if "REG:" in ln][0])
