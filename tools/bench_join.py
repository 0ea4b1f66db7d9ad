#!/usr/bin/env python
"""Equi-JOIN timing (BASELINE config 5 shape, scaled down): orders(id,price,tax,quantity,customer_id)
JOIN customers(id,name,email,since) ON customer_id = id.  usage: python tools/bench_join.py [orders] [customers]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_cases as pc
from cq_b200 import _abi as A
from cq_b200.engine import Table, gpu

L = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
R = int(float(sys.argv[2])) if len(sys.argv) > 2 else 2_000_000
rng = np.random.default_rng(5)


def table_bytes(header, cols):
    # cols: list of numpy arrays of strings (bytes dtype); joined row-wise with ',' and '\n'
    out = cols[0]
    for c in cols[1:]:
        out = np.char.add(np.char.add(out, b","), c)
    return header + b"\n".join(out.tolist()) + b"\n"


t0 = time.time()
oid = np.arange(1, L + 1).astype("S")
price = np.char.mod(b"%.2f", rng.integers(100, 100000, L) / 100)
tax = np.char.mod(b"%.2f", rng.integers(0, 1000, L) / 100)
qty = rng.integers(1, 10, L).astype("S")
cid = rng.integers(1, int(R * 1.1), L).astype("S")
orders = table_bytes(b"id,price,tax,quantity,customer_id\n", [oid, price, tax, qty, cid])
rid = np.arange(1, R + 1).astype("S")
name = np.char.add(b"cust", (np.arange(R) % 9973).astype("S"))
email = np.char.add(np.char.add(b"c", rid), b"@example.com")
since = (2015 + np.arange(R) % 10).astype("S")
customers = table_bytes(b"id,name,email,since\n", [rid, name, email, since])
print(f"generated orders {len(orders)/1e6:.1f} MB ({L} rows), customers {len(customers)/1e6:.1f} MB ({R} rows) in {time.time()-t0:.1f}s")
lib = gpu(); lib.set_device(0)
og = Table.from_bytes(orders, lib=lib); cg = Table.from_bytes(customers, lib=lib)
specs = {
    "count": dict(aggs=[(A.AGG_COUNT_STAR, -1)]),
    "group_since": dict(group_by=[8], out_cols=[8], aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, 1)]),
}
for name_, spec in specs.items():
    for rep in range(2):
        t0 = time.time()
        r = og.execute_raw(pc.build(spec, join=(cg, 4, 0)))
        dt = time.time() - t0
        nb = len(orders) + len(customers)
        print(f"{name_}: groups {r['n_groups']} count0 {r['count0']} kernel_ms {r['kernel_ms']:.2f} wall_ms {dt*1e3:.1f} "
              f"GB/s(wall) {nb/dt/1e9:.2f} rows/s {(L+R)/dt/1e6:.1f}M")
