"""N>1 host logic on CPU: world size 2 over gloo. Each rank runs the (CPU oracle) scan on its own
byte-range shard of one file, the partial group results are exchanged with a collective, merged
(count/sum add, min/max combine, first appearance = smallest global offset) and must equal the
single-shard answer. This is the contract the GPU ranks implement with cqg_partial_* + NCCL
(SURVEY.md §8e); the GPU side of it is measured by `bench.py --gpus N`."""
import os
import socket

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import parity_cases as pc
from cq_b200 import _abi as A
from cq_b200.engine import Table
from oracle_lib import generate_bigdata, oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def merge_partials(parts, aggs):
    """parts: list (per rank) of engine result dicts for the same plan."""
    merged = {}
    for res in parts:
        for g in res["groups"]:
            key = tuple(g["out"])
            m = merged.get(key)
            if m is None:
                merged[key] = {"first_offset": g["first_offset"], "count": g["count"], "out": g["out"],
                               "sum": list(g["sum"]), "ncount": list(g["ncount"]), "aggs": list(g["aggs"])}
                continue
            m["first_offset"] = min(m["first_offset"], g["first_offset"])
            m["count"] += g["count"]
            for a, (f, _) in enumerate(aggs):
                if f in (A.AGG_SUM, A.AGG_AVG):
                    m["sum"][a] += g["sum"][a]
                    m["ncount"][a] += g["ncount"][a]
                elif f in (A.AGG_MIN, A.AGG_MAX):
                    x, y = m["aggs"][a], g["aggs"][a]
                    if x[0] == "N":
                        m["aggs"][a] = y
                    elif y[0] != "N" and x[0] in "ID" and y[0] in "ID":
                        better = y[1] < x[1] if f == A.AGG_MIN else y[1] > x[1]
                        if better:
                            m["aggs"][a] = y
    out = sorted(merged.values(), key=lambda m: m["first_offset"])
    for m in out:
        for a, (f, _) in enumerate(aggs):
            if f in (A.AGG_COUNT_STAR, A.AGG_COUNT):
                m["aggs"][a] = ("I", m["count"])
            elif f == A.AGG_SUM:
                m["aggs"][a] = ("D", m["sum"][a])
            elif f == A.AGG_AVG:
                m["aggs"][a] = ("D", m["sum"][a] / m["ncount"][a] if m["ncount"][a] else 0.0)
    return out


def _worker(rank, world, port, data, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        spec = pc.plans()["group_name"]
        with Table.from_bytes(data, lib=oracle()) as t:
            t.set_shard(rank, world)
            part = t.execute(pc.build(spec))
        gathered = [None] * world
        dist.all_gather_object(gathered, part)
        if rank == 0:
            q.put(gathered)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_shards_merge_to_the_whole_answer():
    data = generate_bigdata(20_000, seed=9)
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, data, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    spec = pc.plans()["group_name"]
    with Table.from_bytes(data, lib=oracle()) as t:
        whole = t.execute(pc.build(spec))
    assert sum(p["rows_scanned"] for p in gathered) == whole["rows_scanned"] == 20_000
    merged = merge_partials(gathered, spec["aggs"])
    assert len(merged) == len(whole["groups"])
    for m, w in zip(merged, whole["groups"]):
        assert m["out"] == w["out"] and m["count"] == w["count"] and m["first_offset"] == w["first_offset"]
        for x, y in zip(m["aggs"], w["aggs"]):
            assert x[0] == y[0]
            if x[0] == "D":
                assert abs(x[1] - y[1]) <= 1e-12 * max(abs(x[1]), abs(y[1]))
            else:
                assert x == y
