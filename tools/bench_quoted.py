#!/usr/bin/env python
"""BASELINE config 4 shape: quoted / escaped-comma CSV with string predicates.
usage: python tools/bench_quoted.py [rows]"""
import os, sys, time, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_cases as pc
from cq_b200 import _abi as A
from cq_b200.engine import Table, gpu
from oracle_lib import oracle

rows = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
rnd = random.Random(4)
first = ["Ada", "Brook", "Cyrus", "Dana", "Eli", "Fay", "Gus", "Hana", "Ivo", "Jude", "Max", "Xena", "Alex"]
last = ["Smith", "Jones", "Lee", "Fox", "Marx", "Nguyen", "O'Neil", "Baxter"]
roles = ["admin", "user", "moderator"]
block = []
for i in range(100_000):
    f, l = rnd.choice(first), rnd.choice(last)
    name = f'"{l}, {f}"' if i % 3 else (f'"say ""{f}"""' if i % 2 else f + l)
    block.append(f"{name},{rnd.choice(roles)},{rnd.randint(10, 80)},{rnd.randint(100, 200) / 100}")
blk = ("\n".join(block) + "\n").encode()
data = b"name,role,age,height\n" + blk * (rows // 100_000)
print(f"{len(data)/1e6:.1f} MB, {rows} rows")
lib = gpu(); lib.set_device(0)
tg = Table.from_bytes(data, lib=lib)
specs = {
    "role = 'admin'": dict(where=("=", ("col", 1), ("const", "admin")), aggs=[(A.AGG_COUNT_STAR, -1)]),
    "name LIKE '%x%'": dict(where=("like", ("col", 0), ("const", "%x%")), aggs=[(A.AGG_COUNT_STAR, -1)]),
    "role='admin' AND name LIKE '%x%' GROUP BY role": dict(where=("and", ("=", ("col", 1), ("const", "admin")), ("like", ("col", 0), ("const", "%x%"))),
                                                          group_by=[1], out_cols=[1], aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_AVG, 2)]),
}
small = b"name,role,age,height\n" + blk
with Table.from_bytes(small, lib=lib) as sg, Table.from_bytes(small, lib=oracle()) as so:
    for name, spec in specs.items():
        pc.compare_results(sg.execute(pc.build(spec)), so.execute(pc.build(spec)))
print("parity vs oracle on one block: ok")
for name, spec in specs.items():
    for rep in range(2):
        r = tg.execute_raw(pc.build(spec))
    print(f"{name}: count0 {r['count0']} kernel_ms {r['kernel_ms']:.2f} GB/s {len(data)/r['kernel_ms']/1e6:.1f}")
