#!/usr/bin/env python
"""The drop-in route on N GPUs of one process: cqg_table_open(path) with CQ_GPUS=N on a page-cached 10 GB file, then the
headline GROUP BY, the COUNT and the 1.8 M-group query. Reports end to end (open + stage + query + close) and resident
(query only) GB/s per N and checks every result against N = 1.
usage: python tools/bench_multi.py [bytes] [N ...]"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import parity_cases as pc
from cq_b200.engine import Table, _check, gpu

nbytes = float(sys.argv[1]) if len(sys.argv) > 1 else 1e10
ns = [int(x) for x in sys.argv[2:]] or [n for n in (1, 2, 4, 8) if n <= torch.cuda.device_count()]
lib = gpu()
lib.set_device(0)
rows = int(nbytes / 29.89)
cap = lib.generate_bigdata_bound(rows, 0) + lib.device_padding()
buf = torch.empty(cap, dtype=torch.uint8, device="cuda")
size = C.c_size_t()
_check(lib, lib.generate_bigdata(buf.data_ptr(), cap - lib.device_padding(), rows, 1, 0, C.byref(size)))
n = size.value
host = buf[:n].cpu().numpy()
del buf
torch.cuda.empty_cache()
path = f"/dev/shm/cq_multi_{os.getpid()}.csv"
with open(path, "wb") as f:
    step = 256 << 20
    for o in range(0, n, step):
        f.write(memoryview(host[o:o + step]))
del host
out = {"bytes": n, "runs": {}}
ref = {}
try:
    for N in ns:
        os.environ["CQ_GPUS"] = str(N)
        legs = {}
        for name in os.environ.get("BM_LEGS", "group_name,count_age_gt_40,group_high_card").split(","):
            plan = pc.build(pc.plans()[name])
            def e2e():
                with Table.open(path, lib=lib) as t:
                    return t.execute_raw(plan)
            e2e()
            t0 = time.perf_counter()
            for _ in range(2):
                r = e2e()
            dt = (time.perf_counter() - t0) / 2
            with Table.open(path, lib=lib) as t:
                gpus = lib.table_gpus(t.handle)
                for _ in range(2):
                    t.execute_raw(plan)
                t1 = time.perf_counter()
                for _ in range(3):
                    r2 = t.execute_raw(plan)
                dq = (time.perf_counter() - t1) / 3
                full = t.execute(plan) if name != "group_high_card" else None
            key = (r["n_groups"], r["count0"], r["rows_scanned"])
            if name not in ref:
                ref[name] = (key, full)
            assert key == ref[name][0], (name, N, key, ref[name][0])
            if full is not None:
                pc.compare_results(full, ref[name][1])
            legs[name] = {"gpus": gpus, "e2e_gbs": n / dt / 1e9, "e2e_ms": dt * 1e3, "resident_gbs": n / dq / 1e9, "resident_ms": dq * 1e3,
                          "scan_kernel_ms_max": r2["kernel_ms"], "groups": r["n_groups"]}
        out["runs"][str(N)] = legs
        print(N, json.dumps(legs), flush=True)
finally:
    os.unlink(path)
print(json.dumps(out))
