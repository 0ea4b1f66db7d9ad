#!/usr/bin/env python
"""What one step of the sharded path costs beside its scan, on ONE GPU (no NCCL): cqg_execute_partial + export + new_like +
merge + finish on a 1/8 slice of the 10 GB file, against cqg_execute on the same slice. usage: python tools/time_partial_step.py [bytes]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import parity_cases as pc
from cq_b200 import _abi as A
from cq_b200.engine import Table, _check, gpu
nbytes = float(sys.argv[1]) if len(sys.argv) > 1 else 1.25e9
lib = gpu(); lib.set_device(0)
rows = int(nbytes / 29.89)
cap = lib.generate_bigdata_bound(rows, 0) + lib.device_padding()
buf = torch.empty(cap, dtype=torch.uint8, device="cuda")
size = C.c_size_t()
_check(lib, lib.generate_bigdata(buf.data_ptr(), cap - lib.device_padding(), rows, 1, 0, C.byref(size)))
t = Table.from_device(buf.data_ptr(), size.value, lib=lib, keep=buf)
for name in ("group_name", "count_age_gt_40"):
    spec = dict(pc.plans()[name]); spec["out_cols"] = []
    plan = pc.build(spec)
    send = torch.zeros(64 * 4096, dtype=torch.uint8, device="cuda")
    def direct():
        return t.execute_raw(plan)
    def partial():
        p = C.c_void_p(); _check(lib, lib.execute_partial(t.handle, C.byref(plan.q), C.byref(p)))
        got = C.c_int64(); _check(lib, lib.partial_export(p, 0, 1, send.data_ptr(), 64, C.byref(got)))
        m = C.c_void_p(); _check(lib, lib.partial_new_like(p, C.byref(m)))
        _check(lib, lib.partial_merge(m, send.data_ptr(), got.value))
        res = C.POINTER(A.Result)(); _check(lib, lib.partial_finish(m, t.handle, C.byref(res)))
        k = lib.partial_kernel_ms(p)
        lib.result_free(res); lib.partial_free(m); lib.partial_free(p)
        return {"kernel_ms": k}
    for label, fn in (("cqg_execute", direct), ("partial pipeline", partial)):
        for _ in range(3): r = fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20): r = fn()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
        print(f"{name:18s} {label:18s} {dt*1e3:7.3f} ms per step, scan kernel {r['kernel_ms']:.3f} ms, beside the scan {dt*1e3 - r['kernel_ms']:.3f} ms")
