#!/usr/bin/env python
"""Run one named plan of tests/parity_cases.py on a generated table (for ncu captures).
usage: python tools/run_plan.py <plan> [bytes] [reps]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import parity_cases as pc
from cq_b200.engine import Table, _check, gpu
name = sys.argv[1]; nbytes = float(sys.argv[2]) if len(sys.argv) > 2 else 2e9; reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
lib = gpu(); lib.set_device(0)
rows = int(nbytes / 29.89)
cap = lib.generate_bigdata_bound(rows, 0) + lib.device_padding()
buf = torch.empty(cap, dtype=torch.uint8, device="cuda")
size = C.c_size_t()
_check(lib, lib.generate_bigdata(buf.data_ptr(), cap - lib.device_padding(), rows, 1, 0, C.byref(size)))
t = Table.from_device(buf.data_ptr(), size.value, lib=lib, keep=buf)
plan = pc.build(pc.plans()[name])
for _ in range(reps):
    r = t.execute_raw(plan)
    print(name, "groups", r["n_groups"], "kernel_ms", round(r["kernel_ms"], 3), "GB/s", round(size.value / r["kernel_ms"] / 1e6, 1), "launches", r["kernel_launches"])
