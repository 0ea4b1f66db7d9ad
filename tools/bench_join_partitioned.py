#!/usr/bin/env python
"""Hash-partitioned equi-JOIN over N GPUs (BASELINE config 5 shape: orders JOIN customers ON customer_id = id).
Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
             tools/bench_join_partitioned.py [orders] [customers]
(or plainly `python tools/bench_join_partitioned.py ...` for N = 1). Every rank holds both files whole; row
offsets are split by key owner and exchanged with NCCL all-to-all (cq_b200/partitioned_join.py). Prints one JSON
line on rank 0: wall time per query (max over ranks), the exchange volume, and the single-GPU join on rank 0
for comparison. The result must equal the single-GPU join of the same files."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
import parity_cases as pc
from cq_b200 import _abi as A
from cq_b200 import partitioned_join as pj
from cq_b200.engine import Table, gpu

ARGS = [a for a in sys.argv[1:] if not a.startswith("--")]
DEVICE_GEN = "--device-gen" in sys.argv  # both sides from the seeded device generator (`...,uid`), joined on uid: any size in seconds
L = int(float(ARGS[0])) if len(ARGS) > 0 else 20_000_000
R = int(float(ARGS[1])) if len(ARGS) > 1 else 2_000_000
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rng = np.random.default_rng(5)


def table_bytes(header, cols):
    out = cols[0]
    for c in cols[1:]:
        out = np.char.add(np.char.add(out, b","), c)
    return header + b"\n".join(out.tolist()) + b"\n"


def device_table(lib, rows, seed):
    import ctypes as C
    from cq_b200.engine import _check
    cap = lib.generate_bigdata_bound(rows, R) + lib.device_padding()
    b = torch.empty(cap, dtype=torch.uint8, device="cuda")
    sz = C.c_size_t()
    _check(lib, lib.generate_bigdata(b.data_ptr(), cap - lib.device_padding(), rows, seed, R, C.byref(sz)))
    return Table.from_device(b.data_ptr(), sz.value, lib=lib, keep=b), sz.value


lib = gpu(); lib.set_device(local)
if DEVICE_GEN:
    # `name,surname,age,gender,height,uid` on both sides, uid ~ U{0..R-1}: the 100 M x 10 M shape of BASELINE configs[4]
    og, nl = device_table(lib, L, 11)
    cg, nr = device_table(lib, R, 12)
    KEY_L = KEY_R = 5
    specs = {
        "count": dict(aggs=[(A.AGG_COUNT_STAR, -1)]),
        "group_right_gender_sum_left_age": dict(group_by=[6 + 3], out_cols=[6 + 3], aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, 2)]),
    }
    total_bytes = nl + nr
else:
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
    og = Table.from_bytes(orders, lib=lib); cg = Table.from_bytes(customers, lib=lib)
    KEY_L, KEY_R = 4, 0
    specs = {
        "count": dict(aggs=[(A.AGG_COUNT_STAR, -1)]),
        "group_since": dict(group_by=[8], out_cols=[8], aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, 1)]),
    }
    total_bytes = len(orders) + len(customers)
out = {"orders_rows": L, "customers_rows": R, "bytes": total_bytes, "n_gpus": world, "device_generated": DEVICE_GEN, "queries": {}}
for name_, spec in specs.items():
    plan = pc.build(spec, join=(cg, KEY_L, KEY_R))
    best = None
    for rep in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if world > 1:
            r = pj.join_aggregate(lib, og, cg, plan, dist=dist)
        else:
            r = pj.join_aggregate(lib, og, cg, plan, world=1)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        best = dt if best is None else min(best, dt)
    entry = {"wall_ms": best * 1e3, "gbs": out["bytes"] / best / 1e9, "rows_per_s": (L + R) / best,
             "groups": len(r["groups"]), "count0": r["groups"][0]["count"] if r["groups"] else 0,
             "probe_kernel_ms_this_rank": r["stats"]["kernel_ms"]}
    if rank == 0:
        t0 = time.perf_counter()
        single = og.execute(plan)
        torch.cuda.synchronize()
        entry["single_gpu_wall_ms"] = (time.perf_counter() - t0) * 1e3
        same = len(single["groups"]) == len(r["groups"]) and all(
            a["count"] == b["count"] and a["out"] == b["out"] and a["first_offset"] == b["first_offset"] and
            all(abs(x - y) <= 1e-12 * max(abs(x), abs(y), 1.0) for x, y in zip(a["sum"], b["sum"]))
            for a, b in zip(single["groups"], r["groups"]))
        entry["equals_single_gpu_join"] = bool(same)
    out["queries"][name_] = entry
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
