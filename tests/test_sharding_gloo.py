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


# ---------------------------------------------------------------------------------------------
# hash-partitioned join: the exchange the GPU ranks do with cqg_partition_rows + NCCL all-to-all
# ---------------------------------------------------------------------------------------------
def _key_owner(text, world):
    """owner of a join key: equal under value_compare (src/csv_reader.c:98-130) => same owner"""
    t = text.strip()
    try:
        k = ("n", float(t))
    except ValueError:
        k = ("s", t)
    if t == "":
        k = ("null",)
    import zlib
    return zlib.crc32(repr(k).encode()) % world


def _shard_rows(data, rank, world):
    """data rows (bytes, without the header) whose first byte lies in this rank's byte range"""
    lo, hi = len(data) * rank // world, len(data) * (rank + 1) // world
    return [ln for ln, p in _rows_with_pos(data) if (lo <= p or rank == 0) and p < hi]


def _rows_with_pos(data):
    header_end = data.index(b"\n") + 1
    pos = header_end
    for line in data[header_end:].split(b"\n"):
        if line:
            yield line, pos
        pos += len(line) + 1


def _join_worker(rank, world, port, orders, customers, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oh, ch = orders[:orders.index(b"\n") + 1], customers[:customers.index(b"\n") + 1]
        # partition this rank's shard of both sides by key owner
        send_l = [[] for _ in range(world)]
        send_r = [[] for _ in range(world)]
        for line in _shard_rows(orders, rank, world):
            send_l[_key_owner(line.split(b",")[4].decode(), world)].append(line)
        for line in _shard_rows(customers, rank, world):
            send_r[_key_owner(line.split(b",")[0].decode(), world)].append(line)
        # all-to-all (gloo: as an all-gather of the send lists, each rank keeps its own column)
        all_l, all_r = [None] * world, [None] * world
        dist.all_gather_object(all_l, send_l)
        dist.all_gather_object(all_r, send_r)
        mine_l = [ln for src in all_l for ln in src[rank]]
        mine_r = [ln for src in all_r for ln in src[rank]]
        # local build + probe + aggregate over the owned rows
        spec = dict(group_by=[8], out_cols=[8], aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, 1)])
        with Table.from_bytes(oh + b"\n".join(mine_l) + b"\n", lib=oracle()) as lt, \
                Table.from_bytes(ch + b"\n".join(mine_r) + b"\n", lib=oracle()) as rt:
            part = lt.execute(pc.build(spec, join=(rt, 4, 0)))
        gathered = [None] * world
        dist.all_gather_object(gathered, (part, len(mine_l), len(mine_r)))
        if rank == 0:
            q.put(gathered)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_hash_partitioned_join_equals_the_whole_join():
    import random
    rnd = random.Random(3)
    orders = ["id,price,tax,quantity,customer_id"]
    for i in range(3000):
        key = rnd.choice([str(rnd.randint(1, 450)), f"{rnd.randint(1, 450)}.0", "", f"00{rnd.randint(1, 9)}"])
        orders.append(f"{i + 1},{rnd.randint(100, 9999) / 100:.2f},0.5,{rnd.randint(1, 9)},{key}")
    customers = ["id,name,email,since"]
    for i in range(400):
        customers.append(f"{i + 1},c{i % 7},e{i}@x.org,{2015 + i % 5}")
    customers += ["7.0,dup,d@x.org,1999", ",nokey,n@x.org,1998"]
    orders = ("\n".join(orders) + "\n").encode()
    customers = ("\n".join(customers) + "\n").encode()
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_join_worker, args=(r, world, port, orders, customers, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    spec = dict(group_by=[8], out_cols=[8], aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, 1)])
    with Table.from_bytes(orders, lib=oracle()) as lt, Table.from_bytes(customers, lib=oracle()) as rt:
        whole = lt.execute(pc.build(spec, join=(rt, 4, 0)))
    assert sum(g[1] for g in gathered) == 3000 and sum(g[2] for g in gathered) == 402  # every row on exactly one rank
    got = {}
    for part, _, _ in gathered:
        for g in part["groups"]:
            m = got.setdefault(tuple(g["out"]), [0, 0.0])
            m[0] += g["count"]
            m[1] += g["sum"][1]
    want = {tuple(g["out"]): (g["count"], g["sum"][1]) for g in whole["groups"]}
    assert set(got) == set(want)
    for k, (c, s) in want.items():
        assert got[k][0] == c
        assert abs(got[k][1] - s) <= 1e-12 * max(abs(s), 1.0)
