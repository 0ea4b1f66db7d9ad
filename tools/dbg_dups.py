import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import importlib.util
spec = importlib.util.spec_from_file_location("t", "tests/test_gpu_plan_parity.py"); m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
import parity_cases as pc
from cq_b200 import _abi as A
from cq_b200.engine import Table, gpu
from oracle_lib import oracle
data = m._lean_stress_table(60000, 21, False)
sp = dict(group_by=[0, 2], out_cols=[0, 2], aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, 3)])
with Table.from_bytes(data, lib=gpu()) as tg, Table.from_bytes(data, lib=oracle()) as to:
    g = tg.execute(pc.build(sp)); w = to.execute(pc.build(sp))
from collections import Counter
cg = Counter(tuple(x["out"]) for x in g["groups"]); cw = Counter(tuple(x["out"]) for x in w["groups"])
print(len(g["groups"]), len(w["groups"]))
d = [k for k, v in cg.items() if v > 1]
print("dups on gpu:", d[:20])
print("only gpu:", [k for k in cg if k not in cw][:10]); print("only oracle:", [k for k in cw if k not in cg][:10])
for k in d[:5]:
    print(k, [ (x["first_offset"], x["count"]) for x in g["groups"] if tuple(x["out"]) == k], [ (x["first_offset"], x["count"]) for x in w["groups"] if tuple(x["out"]) == k])
    off = [x["first_offset"] for x in g["groups"] if tuple(x["out"]) == k]
    for o in off: print("   row:", data[o:o+80].split(b"\n")[0])
