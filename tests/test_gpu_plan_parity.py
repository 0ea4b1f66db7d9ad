"""GPU parity proper: the CUDA path through the C-ABI vs the pinned CPU oracle on the same
seeded inputs and the same plans. Integers, keys, row sets: bit-exact; SUM/AVG: 1e-12 relative."""
import ctypes as C
import random

import pytest

import parity_cases as pc
from cq_b200 import _abi as A
from cq_b200.engine import CqError, Plan, Table, csv_config, gpu
from oracle_lib import generate_bigdata, oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _compile_kernels_per_query_shape():
    """The tables of this file are far below the size at which the library compiles a kernel for a query's shape
    (CQG_JIT_MIN_BYTES, 64 MB): lower the bar so that the plan-level parity tests run those kernels."""
    import os
    old = os.environ.get("CQG_JIT_MIN_BYTES")
    os.environ["CQG_JIT_MIN_BYTES"] = "0"
    yield
    if old is None:
        os.environ.pop("CQG_JIT_MIN_BYTES", None)
    else:
        os.environ["CQG_JIT_MIN_BYTES"] = old


@pytest.fixture(scope="module")
def big():
    data = generate_bigdata(200_000, seed=1)
    return data, Table.from_bytes(data, lib=gpu()), Table.from_bytes(data, lib=oracle())


@pytest.fixture(scope="module")
def big_uid():
    data = generate_bigdata(150_000, seed=5, key_card=40_000)
    return data, Table.from_bytes(data, lib=gpu()), Table.from_bytes(data, lib=oracle())


def test_header_and_row_count(big):
    data, tg, to = big
    assert tg.columns == to.columns == ["name", "surname", "age", "gender", "height"]
    assert tg.row_count() == to.row_count() == 200_000
    assert tg.column_index("HEIGHT") == 4


@pytest.mark.parametrize("name", sorted(pc.plans()))
def test_plan_parity(big, name):
    data, tg, to = big
    spec = pc.plans()[name]
    got = tg.execute(pc.build(spec))
    want = to.execute(pc.build(spec))
    assert got["kernel_launches"] > 0
    pc.compare_results(got, want)


@pytest.mark.parametrize("name", sorted(pc.plans_uid()))
def test_plan_parity_uid(big_uid, name):
    data, tg, to = big_uid
    spec = pc.plans_uid()[name]
    pc.compare_results(tg.execute(pc.build(spec)), to.execute(pc.build(spec)))


@pytest.mark.parametrize("nshards", [2, 3, 8])
def test_byte_range_shards_partition_the_rows(big, nshards):
    """§8e: shard i owns the rows whose first byte lies in its byte range; shards are disjoint and complete."""
    data, tg, to = big
    spec = pc.plans()["group_name"]
    total_rows = 0
    counts = {}
    for i in range(nshards):
        tg.set_shard(i, nshards)
        to.set_shard(i, nshards)
        try:
            got = tg.execute(pc.build(spec))
            want = to.execute(pc.build(spec))
        finally:
            tg.set_shard(0, 1)
            to.set_shard(0, 1)
        pc.compare_results(got, want)
        total_rows += got["rows_scanned"]
        for g in got["groups"]:
            counts[g["out"][0]] = counts.get(g["out"][0], 0) + g["count"]
    assert total_rows == 200_000
    whole = tg.execute(pc.build(spec))
    assert counts == {g["out"][0]: g["count"] for g in whole["groups"]}


def _edge_files():
    rnd = random.Random(7)
    files = {
        "empty_rows_only": b"a,b\n\n\n\n",
        "header_only": b"a,b",
        "header_nl": b"a,b\n",
        "one_row_no_nl": b"a,b\n1,2",
        "crlf": b"a,b\r\n1,2\r\n3,4\r\n\r\n5,6",
        "cr_only": b"a,b\r1,2\r3,4\r",
        "blank_start": b"\n\r\n  \na,b\n1,2\n",
        "ragged": b"a,b,c\n1\n1,2\n1,2,3\n1,2,3,4\n,,\n,\n",
        "spaces": b"a,b\n 1 , 2 \n\t3\t,\t4\t\n  ,  \n5,   \n",
        "quotes": b'a,b\n"1","x,y"\n"2" junk,"say ""hi"""\n"unterminated,3\n4,"also unterminated\n  "5"  ,ok\n',
        "long_row": b"a,b\n" + b"x" * 5000 + b",7\n1,2\n" + b"y" * 70000 + b",9\n3,4\n",
        "long_quoted": b'a,b\n"' + b"q," * 3000 + b'",11\n5,6\n',
        "wide": (",".join(f"c{i}" for i in range(120)) + "\n" + "\n".join(",".join(str(r * 1000 + i) for i in range(120))
                                                                          for r in range(50)) + "\n").encode(),
        "tiny_rows": b"a\n" + b"\n".join(str(i % 10).encode() for i in range(40000)) + b"\n",
        "nul_bytes": b"a,b\n1\x002,x\x00y\n3,4\n",
        "high_bytes": "a,b\né,ü\n1,2\n".encode("utf-8"),
    }
    # rows straddling tile edges at every alignment
    rows = [b"k,v"]
    for i in range(30000):
        rows.append(b"%d,%s" % (rnd.randint(0, 50), b"z" * rnd.randint(0, 40)))
    files["straddle"] = b"\n".join(rows) + b"\n"
    return files


@pytest.mark.parametrize("name", sorted(_edge_files()))
def test_edge_files(name):
    data = _edge_files()[name]
    specs = [
        dict(aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, 0), (A.AGG_MIN, 1), (A.AGG_MAX, 1)]),
        dict(group_by=[0], out_cols=[0, 1], aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, 1), (A.AGG_MAX, 0)]),
        dict(group_by=[1, 0], out_cols=[1], aggs=[(A.AGG_COUNT, 0)]),
        dict(mode="select", where=(">", ("col", 0), ("const", 1)), out_cols=[0, 1, 2]),
        dict(mode="select", out_cols=[1, 0]),
    ]
    if name == "wide":
        specs.append(dict(mode="select", where=("=", ("col", 119), ("const", 3119)), out_cols=[0, 60, 119]))
    with Table.from_bytes(data, lib=gpu()) as tg, Table.from_bytes(data, lib=oracle()) as to:
        assert tg.columns == to.columns
        assert tg.row_count() == to.row_count()
        for spec in specs:
            pc.compare_results(tg.execute(pc.build(spec)), to.execute(pc.build(spec)))
        for n in (2, 5):
            for i in range(n):
                tg.set_shard(i, n)
                to.set_shard(i, n)
                pc.compare_results(tg.execute(pc.build(specs[1])), to.execute(pc.build(specs[1])))
            tg.set_shard(0, 1)
            to.set_shard(0, 1)


def test_no_header_and_delimiters():
    for data, cfg in [(b"1;a;2.5\n2;b;3\n3;a;\n", csv_config(";", '"', False)),
                      (b"id|v\n1|x\n2|'p|q'\n", csv_config("|", "'", True)),
                      (b"id\tv\n1\tx\n2\t\ty\n", csv_config("\t", '"', True))]:
        with Table.from_bytes(data, cfg, lib=gpu()) as tg, Table.from_bytes(data, cfg, lib=oracle()) as to:
            assert tg.columns == to.columns
            for spec in [dict(group_by=[1], out_cols=[1, 0], aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, 2)]),
                         dict(mode="select", out_cols=[0, 1, 2])]:
                pc.compare_results(tg.execute(pc.build(spec)), to.execute(pc.build(spec)))


def _join_tables(n_orders, n_customers, seed):
    rnd = random.Random(seed)
    orders = ["id,price,tax,quantity,customer_id"]
    for i in range(n_orders):
        cid = rnd.randint(1, int(n_customers * 1.2))
        orders.append(f"{i + 1},{rnd.randint(100, 99999) / 100:.2f},{rnd.randint(0, 999) / 100:.2f},{rnd.randint(1, 9)},{cid}")
    customers = ["id,name,email,since"]
    for i in range(n_customers):
        customers.append(f"{i + 1},cust{i % 97},c{i}@example.com,{2015 + i % 10}")
    # duplicates on the right: one left row joins several right rows, in right-file order
    for i in range(0, n_customers, 50):
        customers.append(f"{i + 1},dup{i},d{i}@example.com,{2000 + i % 7}")
    return ("\n".join(orders) + "\n").encode(), ("\n".join(customers) + "\n").encode()


def test_join_parity():
    od, cd = _join_tables(20000, 3000, 11)
    lib_g, lib_o = gpu(), oracle()
    with Table.from_bytes(od, lib=lib_g) as og, Table.from_bytes(cd, lib=lib_g) as cg, \
            Table.from_bytes(od, lib=lib_o) as oo, Table.from_bytes(cd, lib=lib_o) as co:
        specs = [
            dict(aggs=[(A.AGG_COUNT_STAR, -1)]),
            dict(aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, 1), (A.AGG_MIN, 6)], where=(">", ("col", 1), ("const", 500))),
            dict(group_by=[8], out_cols=[8, 6, 0], aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, 1), (A.AGG_MAX, 7)]),
            dict(group_by=[6], out_cols=[6], aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_AVG, 3)],
                 where=("like", ("col", 6), ("const", "cust1%"))),
            dict(mode="select", where=("and", (">", ("col", 1), ("const", 990)), ("<", ("col", 8), ("const", 2016))),
                 out_cols=[0, 4, 5, 6, 8]),
            dict(mode="select", out_cols=[0, 6], max_rows=40),
        ]
        for spec in specs:
            got = og.execute(pc.build(spec, join=(cg, 4, 0)))
            want = oo.execute(pc.build(spec, join=(co, 4, 0)))
            pc.compare_results(got, want)
        # unresolved key columns never match (evaluator_joins.c:54)
        got = og.execute(pc.build(specs[0], join=(cg, -1, 0)))
        assert got["groups"][0]["count"] == 0


@pytest.mark.parametrize("jtype", [A.JOIN_LEFT, A.JOIN_RIGHT, A.JOIN_FULL])
def test_outer_join_parity(jtype):
    """LEFT / RIGHT / FULL joins (evaluator_joins.c:128-171): left rows without a match once with NULL right columns, in
    place; right rows without a match behind everything, in right-file order, with NULL left columns."""
    od, cd = _join_tables(20000, 3000, 23)
    # keys without a partner on either side, NULL keys on both sides (NULL = NULL matches), spellings the reference equates
    od += b"90001,1.00,0.10,1,\n90002,2.00,0.20,2,7.0\n90003,3.00,0.30,3,0007\n90004,4.00,0.40,4,999999\n"
    cd += b",nokey,none@example.com,1999\n7.00,seven,s@example.com,1998\n888888,lonely,l@example.com,1997\n888889,lonely2,l2@example.com,1997\n"
    lib_g, lib_o = gpu(), oracle()
    with Table.from_bytes(od, lib=lib_g) as og, Table.from_bytes(cd, lib=lib_g) as cg, \
            Table.from_bytes(od, lib=lib_o) as oo, Table.from_bytes(cd, lib=lib_o) as co:
        specs = [
            dict(aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_COUNT, 6), (A.AGG_SUM, 1), (A.AGG_MIN, 6), (A.AGG_MAX, 0)]),
            dict(group_by=[8], out_cols=[8, 6, 0], aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, 1), (A.AGG_MAX, 7)]),
            dict(group_by=[4], out_cols=[4, 5, 6], aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_MIN, 0), (A.AGG_MIN, 8)]),
            dict(group_by=[6], out_cols=[6], aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_AVG, 3)],
                 where=("or", ("like", ("col", 6), ("const", "cust1%")), (">", ("col", 1), ("const", 900)))),
            dict(mode="select", where=("or", (">", ("col", 1), ("const", 990)), (">", ("col", 5), ("const", 2990))),
                 out_cols=[0, 4, 5, 6, 8]),
            dict(mode="select", out_cols=[0, 6], max_rows=40),
            dict(mode="select", where=("=", ("col", 8), ("const", 1997)), out_cols=[0, 1, 5, 6]),
        ]
        for spec in specs:
            got = og.execute(pc.build(spec, join=(cg, 4, 0, jtype)))
            want = oo.execute(pc.build(spec, join=(co, 4, 0, jtype)))
            pc.compare_results(got, want)
        # an unresolved key column matches nothing: LEFT keeps every left row once
        if jtype == A.JOIN_LEFT:
            got = og.execute(pc.build(specs[0], join=(cg, -1, 0, jtype)))
            want = oo.execute(pc.build(specs[0], join=(co, -1, 0, jtype)))
            pc.compare_results(got, want)
            assert got["groups"][0]["count"] == og.row_count()
        else:
            with pytest.raises(CqError) as ei:
                og.execute(pc.build(specs[0], join=(cg, -1, 0, jtype)))
            assert ei.value.code == A.ERR_UNSUPPORTED_PLAN


def test_parse_value_matches_oracle():
    lib_g, lib_o = gpu(), oracle()
    samples = [b"", b"1", b"007", b"-5", b"+8", b"3.25", b".5", b"5.", b"-0.0", b"1e5", b"-", b"abc", b" 12 ", b"12 3",
               b"20240115", b"2010010100", b"2023-1-5x", b"555-0001", b"2024-01-15", b"01/02/2024", b"31/12/2023",
               b"2023-02-29", b"9223372036854775807", b"9223372036854775808", b"-9223372036854775809", b"0.1",
               b"0.30000000000000004", b"123456789.123456789", b"1234567.1234567", b"  padded  ", b"+20240115",
               b"1234567890123456789", b"12345678.9", b"1.7976931348623157", b"4.35", b"0.000001", b"8.41"]
    rnd = random.Random(3)
    for _ in range(300):
        k = rnd.randint(1, 18)
        s = "".join(rnd.choice("0123456789") for _ in range(k))
        if rnd.random() < 0.7:
            p = rnd.randint(0, len(s))
            s = s[:p] + "." + s[p:]
        if rnd.random() < 0.3:
            s = "-" + s
        samples.append(s.encode())
    for s in samples:
        a, b = A.Value(), A.Value()
        assert lib_g.parse_value(s, len(s), C.byref(a)) == 0, lib_g.last_error()
        assert lib_o.parse_value(s, len(s), C.byref(b)) == 0
        from cq_b200.engine import py_value
        va, vb = py_value(a), py_value(b)
        if va[0] == "D":
            assert vb[0] == "D" and (va[1].hex() == vb[1].hex()), (s, va, vb)
        else:
            assert va == vb, (s, va, vb)
        lib_g.value_release(C.byref(a))
        lib_o.value_release(C.byref(b))


def test_unsupported_inputs_fail_loudly():
    data = b"k,v\n1,99999999999999999999999999999999999999999999.5\n"
    with Table.from_bytes(data, lib=gpu()) as tg:
        try:
            r = tg.execute(Plan(aggs=[(A.AGG_SUM, 1)]))
        except CqError as e:
            assert e.code == A.ERR_UNSUPPORTED
        else:
            with Table.from_bytes(data, lib=oracle()) as to:
                pc.compare_results(r, to.execute(Plan(aggs=[(A.AGG_SUM, 1)])))


@pytest.mark.parametrize("name", ["group_name", "group_gender_minmax", "group_name_surname", "scalar_aggs", "group_height",
                                  "lean_group_two_keys", "count_age_gt_40"])
@pytest.mark.parametrize("world", [2, 3])
def test_partial_aggregates_merge(big, name, world):
    """The multi-GPU exchange on one device: every "rank" scans its byte-range shard into a partial
    (cqg_execute_partial), the records are exported (all of them, or split by owner = hash % world as
    an all-to-all would), merged into a fresh table (cqg_partial_merge) and finished. Must equal the
    single scan of the whole file."""
    import torch
    data, tg, to = big
    lib = gpu()
    spec = pc.plans()[name]
    plan = pc.build(spec)
    want = to.execute(pc.build(spec))
    parts = []
    try:
        for r in range(world):
            tg.set_shard(r, world)
            p = C.c_void_p()
            assert lib.execute_partial(tg.handle, C.byref(plan.q), C.byref(p)) == 0, lib.last_error()
            parts.append(p)
        tg.set_shard(0, 1)
        rec = lib.partial_record_size(parts[0])
        merged = C.c_void_p()
        assert lib.partial_new_like(parts[0], C.byref(merged)) == 0, lib.last_error()
        for p in parts:
            n = lib.partial_count(p)
            counts = (C.c_int64 * world)()
            assert lib.partial_owner_counts(p, world, counts) == 0
            assert sum(counts) == n
            for owner in range(world):  # owner-filtered export, as the all-to-all of many groups does
                buf = torch.zeros(max(counts[owner], 1) * rec, dtype=torch.uint8, device="cuda")
                got = C.c_int64()
                assert lib.partial_export(p, owner, world, buf.data_ptr(), counts[owner], C.byref(got)) == 0, lib.last_error()
                assert got.value == counts[owner]
                assert lib.partial_merge(merged, buf.data_ptr(), got.value) == 0, lib.last_error()
        res = C.POINTER(A.Result)()
        assert lib.partial_finish(merged, tg.handle, C.byref(res)) == 0, lib.last_error()
        from cq_b200.engine import decode_result
        got = decode_result(res.contents, plan)
        lib.result_free(res)
        lib.partial_free(merged)
        got["rows_scanned"] = sum(lib.partial_rows_scanned(p) for p in parts)
        pc.compare_results(got, want)
    finally:
        tg.set_shard(0, 1)
        for p in parts:
            lib.partial_free(p)


def _lean_stress_table(n, seed, clean):
    """Rows of mixed width (under 32, 32..63, 64 and more bytes) and mixed value shapes in the columns the lean
    kernel decodes: short and 5..7 digit decimals, up to 3 and more fraction digits, signed numbers, dates,
    empty fields, long and short texts. `clean=False` sprinkles blanks / CR / quotes so that tiles get handed
    over to the general kernel."""
    rnd = random.Random(seed)
    words = ["a", "bb", "ccc", "delta", "echo-echo", "f" * 16, "g" * 17, "h" * 40, "NULL", "x1", "k_9"]
    rows = ["k1,k2,n1,n2,pad,d1"]
    for i in range(n):
        k1 = rnd.choice(words[:6] if i % 7 else words)
        k2 = rnd.choice(["u", "v", "w", "12", "1.5", "", "007"])
        r = rnd.random()
        if r < 0.55:
            n1 = str(rnd.randint(0, 9999))
        elif r < 0.75:
            n1 = str(rnd.randint(10000, 9999999))
        elif r < 0.85:
            n1 = f"{rnd.randint(0, 999)}.{rnd.randint(0, 999):03d}"
        elif r < 0.9:
            n1 = str(-rnd.randint(1, 500))
        elif r < 0.93:
            n1 = ""
        elif r < 0.96:
            n1 = "2024-01-15"
        else:
            n1 = f"{rnd.randint(0, 99)}.{rnd.randint(0, 99999):05d}"
        n2 = rnd.choice(["1", "2.5", "10", "0.125", "3.", ".5", "99999", "1234567", "12345678", "abc", ""])
        pad = "p" * rnd.choice([0, 0, 0, 3, 9, 20, 45, 70])
        d1 = rnd.choice(["2.0", "1.25", "7", "", "1e3"])
        if not clean and i % 53 == 0:
            k1 = '"q,uoted"'
        if not clean and i % 31 == 0:
            n1 = " " + n1 + " "
        rows.append(f"{k1},{k2},{n1},{n2},{pad},{d1}")
    text = ("\r\n" if not clean else "\n").join(rows) + "\n"
    return text.encode()


@pytest.mark.parametrize("clean", [True, False])
def test_lean_kernel_hand_over_paths(clean):
    data = _lean_stress_table(60_000, 21, clean)
    K1, K2, N1, N2, PAD, D1 = range(6)
    specs = [
        dict(where=(">", ("col", N1), ("const", 5000)), aggs=[(A.AGG_COUNT_STAR, -1)]),
        dict(where=("<=", ("col", N1), ("const", 123.5)), aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, N2), (A.AGG_AVG, N1)]),
        dict(where=("and", ("!=", ("col", K2), ("const", "u")), (">=", ("col", N2), ("const", 1))),
             aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, D1), (A.AGG_COUNT, K1)]),
        dict(where=("or", ("=", ("col", K1), ("const", "delta")), ("<", ("col", N1), ("const", 10))),
             aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_AVG, N2)]),
        dict(group_by=[K2], out_cols=[K2], aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, N1), (A.AGG_AVG, N2)]),
        dict(where=("not", ("=", ("col", N2), ("const", 2.5))), group_by=[K1, K2], out_cols=[K1, K2],
             aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, N2), (A.AGG_MIN, N1), (A.AGG_MAX, N1)]),
        dict(group_by=[N2], out_cols=[N2], aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_MAX, D1), (A.AGG_MIN, N2)]),
        dict(group_by=[K1, N1], out_cols=[K1, N1], aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, N2)]),  # many groups: global mode
        dict(group_by=[D1, K2, K1], out_cols=[D1], aggs=[(A.AGG_COUNT, N1), (A.AGG_AVG, N1), (A.AGG_MAX, N2)]),
    ]
    with Table.from_bytes(data, lib=gpu()) as tg, Table.from_bytes(data, lib=oracle()) as to:
        for spec in specs:
            pc.compare_results(tg.execute(pc.build(spec)), to.execute(pc.build(spec)), rel=1e-11)
        for i in range(3):
            tg.set_shard(i, 3)
            to.set_shard(i, 3)
            for spec in (specs[1], specs[5]):
                pc.compare_results(tg.execute(pc.build(spec)), to.execute(pc.build(spec)), rel=1e-11)


def test_partial_aggregates_merge_with_join():
    """Sharded left table, replicated build side: two partials merged must equal the single join."""
    import torch
    od, cd = _join_tables(20000, 3000, 13)
    lib, lo = gpu(), oracle()
    spec = dict(group_by=[8], out_cols=[8, 6], aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, 1), (A.AGG_MAX, 3)],
                where=(">", ("col", 1), ("const", 100)))
    with Table.from_bytes(od, lib=lib) as og, Table.from_bytes(cd, lib=lib) as cg, \
            Table.from_bytes(od, lib=lo) as oo, Table.from_bytes(cd, lib=lo) as co:
        want = oo.execute(pc.build(spec, join=(co, 4, 0)))
        plan = pc.build(spec, join=(cg, 4, 0))
        parts = []
        world = 2
        for r in range(world):
            og.set_shard(r, world)
            p = C.c_void_p()
            assert lib.execute_partial(og.handle, C.byref(plan.q), C.byref(p)) == 0, lib.last_error()
            parts.append(p)
        og.set_shard(0, 1)
        rec = lib.partial_record_size(parts[0])
        merged = C.c_void_p()
        assert lib.partial_new_like(parts[0], C.byref(merged)) == 0
        for p in parts:
            n = lib.partial_count(p)
            buf = torch.zeros(max(n, 1) * rec, dtype=torch.uint8, device="cuda")
            got = C.c_int64()
            assert lib.partial_export(p, 0, 1, buf.data_ptr(), n, C.byref(got)) == 0
            assert lib.partial_merge(merged, buf.data_ptr(), got.value) == 0
        res = C.POINTER(A.Result)()
        assert lib.partial_finish(merged, og.handle, C.byref(res)) == 0, lib.last_error()
        from cq_b200.engine import decode_result
        got = decode_result(res.contents, plan)
        lib.result_free(res)
        got["rows_scanned"] = sum(lib.partial_rows_scanned(p) for p in parts)
        pc.compare_results(got, want)
        lib.partial_free(merged)
        for p in parts:
            lib.partial_free(p)


def _sparse_anomaly_table(n, seed, kind):
    """A table the scalar lean kernel covers, with ONE kind of byte it does not classify sprinkled in every few
    hundred rows: the tile holding it must come out exactly as the reference reads it (handed to the general
    kernel whole), its neighbours untouched."""
    rnd = random.Random(seed)
    rows = ["name,age,score,tag"]
    for i in range(n):
        name = rnd.choice(["ann", "bob", "carla", "dmitri", "eve"])
        age = str(rnd.randint(0, 120))
        score = rnd.choice(["1.5", "2", "0.25", "10.5", "99", "7.", ".5", "100", "1234", "12345", "3.125"])
        tag = rnd.choice(["x", "yy", "zzz"])
        if i % 397 == 5:
            if kind == "blank_in_field":
                name = "new york"
            elif kind == "blank_number":
                age = " " + age
            elif kind == "quoted":
                name = '"o,k"'
            elif kind == "tab":
                tag = "t\tt"
            elif kind == "empty_line":
                rows.append("")
            elif kind == "cr":
                tag = tag + "\r"
            elif kind == "ragged":
                rows.append(name)
                continue
            elif kind == "long_row":
                tag = "L" * rnd.choice([40, 64, 200, 1500])
            elif kind == "signed":
                age = "-" + age
            elif kind == "bang":
                tag = "!" + tag
        rows.append(f"{name},{age},{score},{tag}")
    return ("\n".join(rows) + "\n").encode()


@pytest.mark.parametrize("kind", ["none", "blank_in_field", "blank_number", "quoted", "tab", "empty_line", "cr", "ragged",
                                  "long_row", "signed", "bang"])
def test_scalar_lean_kernel_dirty_tiles(kind):
    data = _sparse_anomaly_table(50_000, 3, kind)
    NAME_, AGE_, SCORE_, TAG_ = range(4)
    specs = [
        dict(where=(">", ("col", AGE_), ("const", 40)), aggs=[(A.AGG_COUNT_STAR, -1)]),
        dict(where=("<=", ("col", SCORE_), ("const", 2.5)), aggs=[(A.AGG_COUNT_STAR, -1)]),
        dict(where=("=", ("col", SCORE_), ("const", 7)), aggs=[(A.AGG_COUNT_STAR, -1)]),
        dict(where=("!=", ("col", AGE_), ("const", 33.5)), aggs=[(A.AGG_COUNT_STAR, -1)]),
        dict(where=("<", ("col", NAME_), ("const", 3)), aggs=[(A.AGG_COUNT_STAR, -1)]),
        dict(where=("and", (">=", ("col", AGE_), ("const", 18)), ("=", ("col", TAG_), ("const", "yy"))),
             aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, SCORE_), (A.AGG_AVG, AGE_), (A.AGG_COUNT, NAME_)]),
        dict(aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, AGE_), (A.AGG_AVG, SCORE_)], out_cols=[NAME_, TAG_]),
        dict(where=("or", ("<", ("col", SCORE_), ("const", 1)), ("not", ("!=", ("col", NAME_), ("const", "eve")))),
             aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, AGE_)]),
    ]
    with Table.from_bytes(data, lib=gpu()) as tg, Table.from_bytes(data, lib=oracle()) as to:
        assert tg.row_count() == to.row_count()
        for spec in specs:
            pc.compare_results(tg.execute(pc.build(spec)), to.execute(pc.build(spec)))
        for i in range(4):
            tg.set_shard(i, 4)
            to.set_shard(i, 4)
            for spec in (specs[0], specs[5]):
                pc.compare_results(tg.execute(pc.build(spec)), to.execute(pc.build(spec)))


@pytest.mark.parametrize("dense", [True, False])
def test_blanks_are_ordinary_bytes_on_the_lean_kernels(dense):
    """Real CSV has blanks (`New York`, `2024-01-01 10:00`): a blank no longer sends a tile to the general kernel
    (kLeanSpecialXor, cqg_lean.cuh). Interior blanks stay in keys and text literals as they are; a field that STARTS or
    ENDS with one is trimmed by the reference (src/csv_reader.c:195-240) and must come out that way (those rows are
    handed over). Every lean kernel: scalar, few groups, few groups with MIN/MAX, many groups."""
    rnd = random.Random(9)
    cities = ["New York", "San Jose", "Rio", "Los Angeles", " Lima", "Oslo ", "  ", "St. John s", "Quito"]
    rows = ["city,zone,age,height"]
    n = 60_000
    for i in range(n):
        city = rnd.choice(cities) if (dense or i % 211 == 3) else rnd.choice(["Rio", "Quito", "Bonn"])
        zone = f"z{rnd.randint(0, 40 if dense else 3000)}"
        age = str(rnd.randint(10, 80))
        if i % 503 == 7:
            age = " " + age
        if i % 509 == 9:
            age = age + " "
        height = f"{rnd.randint(100, 200) / 100}"
        rows.append(f"{city},{zone},{age},{height}")
    data = ("\n".join(rows) + "\n").encode()
    CITY, ZONE, AGE_, H = range(4)
    specs = [
        dict(where=(">", ("col", AGE_), ("const", 40)), aggs=[(A.AGG_COUNT_STAR, -1)]),
        dict(where=("=", ("col", CITY), ("const", "New York")), aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, AGE_)]),
        dict(where=("!=", ("col", CITY), ("const", "Lima")), aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_AVG, H)]),
        dict(where=("=", ("col", CITY), ("const", "Oslo")), aggs=[(A.AGG_COUNT_STAR, -1)]),
        dict(where=(">", ("col", AGE_), ("const", 25)), group_by=[CITY], out_cols=[CITY],
             aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_AVG, H), (A.AGG_SUM, AGE_)]),
        dict(group_by=[CITY], out_cols=[CITY], aggs=[(A.AGG_MIN, H), (A.AGG_MAX, AGE_), (A.AGG_COUNT_STAR, -1)]),
        dict(group_by=[CITY, ZONE], out_cols=[CITY, ZONE], aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, AGE_), (A.AGG_MAX, H)]),
        dict(group_by=[CITY, ZONE, AGE_], out_cols=[CITY, ZONE, AGE_], aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, AGE_)]),
    ]
    with Table.from_bytes(data, lib=gpu()) as tg, Table.from_bytes(data, lib=oracle()) as to:
        assert tg.row_count() == to.row_count() == n
        for spec in specs:
            pc.compare_results(tg.execute(pc.build(spec)), to.execute(pc.build(spec)))


def _few_groups_table(n, seed, keys, odd_every=0):
    """role-like key column (few values) + age / score / price columns; every `odd_every` rows one of the things the
    written-out few-groups loop (cqg_lean2k.cuh) must hand over, count as NULL, or number as a group of its own."""
    rnd = random.Random(seed)
    rows = ["id,role,age,score,price,note"]
    odd = ["", "NULL", " lead", "lead ", "7", "07", "7.0", "2024-01-05", "-3", "seventeen-bytes-xx", "sixteen-bytes-xxx", "new york"]
    for i in range(n):
        role = rnd.choice(keys)
        age = str(rnd.randint(0, 99))
        score = rnd.choice(["1.5", "2", "0.25", "10.5", "99", "7.", ".5", "100", "1234", "3.12"])
        price = rnd.choice(["12345", "99999.9", "1234.56", "1000000", "5", "0.125"])
        note = rnd.choice(["x", "yy", "a longer note that makes the row wide", "z" * 70])
        if odd_every and i % odd_every == 3:
            k = (i // odd_every) % 6
            if k == 0:
                role = odd[(i // odd_every // 6) % len(odd)]
            elif k == 1:
                age = rnd.choice(["", "NULL", "abc", "-5", "1e2"])
            elif k == 2:
                score = rnd.choice(["", "1.2345", "x"])
            elif k == 3:
                price = rnd.choice(["", "12345678", "1.5e3"])
            elif k == 4:
                rows.append(f"{i},{role}")  # ragged
                continue
            else:
                rows.append("")  # empty line
        rows.append(f"{i},{role},{age},{score},{price},{note}")
    return ("\n".join(rows) + "\n").encode()


@pytest.mark.parametrize("case", ["plain", "odd_rows", "seventeen_groups", "forty_groups"])
def test_few_groups_written_out_loop(case):
    """lean2k_kernel (cqg_lean2k.cuh): one text key, COUNT / SUM / AVG, decimal leaves. Keys are grouped by their raw bytes
    per CTA and made canonical once per group: blanks at either end, the text NULL, empty keys, "7" / "07" / "7.0", dates
    and 17-byte keys must come out as the reference groups them; NULL and 5..7-byte operands, wide and ragged rows; a 17th
    group in a CTA moves the scan to lean2g_kernel, a 33rd to the global table."""
    roles = ["admin", "user", "guest", "moderator"]
    if case == "seventeen_groups":
        roles = [f"role{k}" for k in range(17)]
    elif case == "forty_groups":
        roles = [f"r{k}" for k in range(40)]
    data = _few_groups_table(120_000, 11, roles, odd_every=0 if case == "plain" else 97)
    if case in ("seventeen_groups", "forty_groups"):
        # the planner samples the head of the file: keep it to few keys so that the scan STARTS on lean2k_kernel
        head = _few_groups_table(600, 12, ["admin", "user"])
        data = head + data[data.index(b"\n") + 1:]
    ID, ROLE, AGE_, SCORE, PRICE, NOTE = range(6)
    specs = [
        dict(where=(">", ("col", AGE_), ("const", 25)), group_by=[ROLE], out_cols=[ROLE],
             aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_AVG, AGE_)]),
        dict(group_by=[ROLE], out_cols=[ROLE, ID], aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, SCORE), (A.AGG_AVG, PRICE), (A.AGG_SUM, AGE_)]),
        dict(where=("or", ("<", ("col", SCORE), ("const", 1)), ("not", (">=", ("col", AGE_), ("const", 50)))), group_by=[ROLE],
             out_cols=[ROLE], aggs=[(A.AGG_SUM, PRICE), (A.AGG_COUNT, NOTE)]),
        dict(where=("!=", ("col", AGE_), ("const", 33)), group_by=[ROLE], out_cols=[ROLE]),
    ]
    lib = gpu()
    with Table.from_bytes(data, lib=lib) as tg, Table.from_bytes(data, lib=oracle()) as to:
        before = lib.kernel_launches_named(b"lean2k")
        for spec in specs:
            pc.compare_results(tg.execute(pc.build(spec)), to.execute(pc.build(spec)))
        assert lib.kernel_launches_named(b"lean2k") >= before + len(specs)  # (needs libnvrtc: the kernel is compiled per query)
        for i in range(3):
            tg.set_shard(i, 3)
            to.set_shard(i, 3)
            pc.compare_results(tg.execute(pc.build(specs[1])), to.execute(pc.build(specs[1])))


@pytest.mark.parametrize("kind", ["crlf", "mixed", "cr_only"])
def test_cr_lf_files_stay_on_the_lean_kernels(kind):
    """A file whose lines end in CR LF (DevPlan::crlf, guessed from the head of the file): the group-by kernels take '\\r' as
    one more terminator, the scalar kernel takes the pair as the line end; results as csv_load splits such a file
    (src/csv_reader.c:404-427). `mixed`: some lines end in a bare LF (the scalar kernel hands those tiles over), `cr_only`: old
    Mac line ends."""
    data = generate_bigdata(120_000, seed=4)
    lines = data.split(b"\n")
    if kind == "crlf":
        data = b"\r\n".join(lines)
    elif kind == "cr_only":
        data = b"\r".join(lines)
    else:
        data = b"".join(ln + (b"\n" if i % 1000 == 17 else b"\r\n") for i, ln in enumerate(lines[:-1]))
    lib = gpu()
    names = ["count_age_gt_40", "count_height_gt_1_5", "scalar_aggs", "group_name", "lean_group_two_keys", "group_high_card",
             "lean_group_abort_many", "group_gender_minmax"]
    with Table.from_bytes(data, lib=lib) as tg, Table.from_bytes(data, lib=oracle()) as to:
        assert tg.row_count() == to.row_count() == 120_000
        for name in names:
            spec = pc.plans()[name]
            pc.compare_results(tg.execute(pc.build(spec)), to.execute(pc.build(spec)))
            # (plans with MIN / MAX run on the first lean kernel, which still hands CR tiles over)
            if kind == "crlf" and name in ("count_age_gt_40", "group_name", "lean_group_two_keys", "group_high_card"):
                tiles, handed, rows = C.c_int64(), C.c_int64(), C.c_int64()
                lib.last_scan_stats(C.byref(tiles), C.byref(handed), C.byref(rows))
                assert handed.value <= 2, (name, tiles.value, handed.value)  # (the two tiles at the file's edges)
        for i in range(3):
            tg.set_shard(i, 3)
            to.set_shard(i, 3)
            for name in ("count_age_gt_40", "group_name"):
                spec = pc.plans()[name]
                pc.compare_results(tg.execute(pc.build(spec)), to.execute(pc.build(spec)))


@pytest.mark.parametrize("tail", ["newline", "none", "blank_lines", "crlf_none"])
def test_edge_tiles_inside_the_lean_kernels(tail):
    """Scans of >= 64 tiles keep the tiles at the file's edges on the lean kernel (DevPlan::edge_in_kernel, LeanEdge in
    cqg_lean2.cuh): the first tile has nothing in front of it, the last one ends with the file - with or without a final
    line terminator, or after trailing blank lines. Every lean kernel, whole file and shards; no tile may be handed over."""
    data = generate_bigdata(90_000, seed=31)  # ~2.7 MB = 165 tiles
    if tail == "none":
        data = data[:-1]
    elif tail == "blank_lines":
        data = data + b"\n\n\n"
    elif tail == "crlf_none":
        data = b"\r\n".join(data.split(b"\n"))[:-2]
    lib = gpu()
    names = ["count_age_gt_40", "scalar_aggs", "group_name", "lean_group_two_keys", "group_high_card", "count_height_gt_1_5"]
    with Table.from_bytes(data, lib=lib) as tg, Table.from_bytes(data, lib=oracle()) as to:
        assert tg.row_count() == to.row_count() == 90_000
        for name in names:
            spec = pc.plans()[name]
            pc.compare_results(tg.execute(pc.build(spec)), to.execute(pc.build(spec)))
            if name != "scalar_aggs":  # (MIN / MAX plans: the first lean kernel still hands its edge tiles over)
                tiles, handed, rows = C.c_int64(), C.c_int64(), C.c_int64()
                lib.last_scan_stats(C.byref(tiles), C.byref(handed), C.byref(rows))
                # (the scalar kernel hands over a tile that holds an empty line, and in its CR LF mode one whose last line
                # does not end in the pair: one tile at most, for what is IN the tile, not for being at the edge)
                assert tiles.value >= 64 and handed.value <= (1 if tail in ("crlf_none", "blank_lines") else 0), (name, tiles.value, handed.value)
        for i in range(3):
            tg.set_shard(i, 3)
            to.set_shard(i, 3)
            for name in ("count_age_gt_40", "group_name", "group_high_card"):
                spec = pc.plans()[name]
                pc.compare_results(tg.execute(pc.build(spec)), to.execute(pc.build(spec)))


def test_signed_numbers_stay_on_the_lean_kernels():
    """Fields with a leading sign (temperatures, balances): decoded on the lean kernels' cold paths (cqg_lean2.cuh:
    CQG_L2_SIGNED; the sign rides in fd16 and picks the mirrored interval), not handed over row by row - a column of them used
    to abort the whole scan to the general kernel. Values as strtoll / strtod read them: -0, -.5, -5., +7."""
    rnd = random.Random(17)
    temps = ["-12.5", "-3", "+7", "-0", "-0.0", "-.5", "-5.", "12.25", "100", "-99.9", "0.5", "-1", "-123.4", "-1234.56", "+0.25", "-7"]
    rows = ["city,temp,balance,tag"]
    n = 80_000
    for i in range(n):
        city = rnd.choice(["Oslo", "Lima", "Quito", "Bonn", "Riga"])
        temp = rnd.choice(temps[:13]) if i % 101 else rnd.choice(temps)
        bal = str(rnd.randint(-999, 999))
        rows.append(f"{city},{temp},{bal},t{i % 5}")  # (25 groups: the few-groups kernels number up to 32 per CTA)
    data = ("\n".join(rows) + "\n").encode()
    CITY, TEMP, BAL, TAG = range(4)
    specs = [
        dict(where=(">", ("col", TEMP), ("const", -5)), aggs=[(A.AGG_COUNT_STAR, -1)]),
        dict(where=("<", ("col", BAL), ("const", 0)), aggs=[(A.AGG_COUNT_STAR, -1)]),
        dict(where=("=", ("col", TEMP), ("const", 0)), aggs=[(A.AGG_COUNT_STAR, -1)]),
        dict(where=("!=", ("col", BAL), ("const", -7)), aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, BAL)]),
        dict(where=("and", (">=", ("col", BAL), ("const", -100)), ("<=", ("col", TEMP), ("const", -0.5))),
             aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, TEMP), (A.AGG_AVG, BAL)]),
        dict(aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, TEMP), (A.AGG_AVG, TEMP), (A.AGG_SUM, BAL)]),
        dict(where=(">=", ("col", BAL), ("const", -100)), group_by=[CITY], out_cols=[CITY],
             aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, TEMP), (A.AGG_AVG, BAL)]),
        dict(where=("<", ("col", TEMP), ("const", 0)), group_by=[CITY, TAG], out_cols=[CITY, TAG],
             aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, BAL), (A.AGG_AVG, TEMP)]),
    ]
    lib = gpu()
    with Table.from_bytes(data, lib=lib) as tg, Table.from_bytes(data, lib=oracle()) as to:
        assert tg.row_count() == to.row_count() == n
        for k, spec in enumerate(specs):
            pc.compare_results(tg.execute(pc.build(spec)), to.execute(pc.build(spec)))
            tiles, handed, rows_h = C.c_int64(), C.c_int64(), C.c_int64()
            lib.last_scan_stats(C.byref(tiles), C.byref(handed), C.byref(rows_h))
            # only the 8-byte `-1234.56` (one row in 1600) is outside the lean repertoire
            assert handed.value <= 2 and rows_h.value < n // 500, (k, tiles.value, handed.value, rows_h.value)  # (tiles: the file's edges)
        for i in range(3):
            tg.set_shard(i, 3)
            to.set_shard(i, 3)
            for spec in (specs[0], specs[4], specs[6]):
                pc.compare_results(tg.execute(pc.build(spec)), to.execute(pc.build(spec)))


@pytest.mark.parametrize("world", [2, 5])
def test_hash_partitioned_join(world):
    """BASELINE config 5 on one device: `world` simulated ranks split the row offsets of their shards by key
    owner (cqg_partition_rows), the lists are regrouped as the NCCL all-to-all regroups them, every owner builds
    and probes over its own rows (cqg_execute_partial_rows), the partials merge. Must equal the single join."""
    from cq_b200 import partitioned_join as pj
    od, cd = _join_tables(20000, 3000, 17)
    # keys the reference equates across spellings, NULL keys on both sides, keys without a partner
    od += b"90001,1.00,0.10,1,\n90002,2.00,0.20,2,7.0\n90003,3.00,0.30,3,0007\n90004,4.00,0.40,4,999999\n"
    cd += b",nokey,none@example.com,1999\n7.00,seven,s@example.com,1998\n"
    lib_g, lib_o = gpu(), oracle()
    with Table.from_bytes(od, lib=lib_g) as og, Table.from_bytes(cd, lib=lib_g) as cg, \
            Table.from_bytes(od, lib=lib_o) as oo, Table.from_bytes(cd, lib=lib_o) as co:
        specs = [
            dict(aggs=[(A.AGG_COUNT_STAR, -1)]),
            dict(aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, 1), (A.AGG_MIN, 6), (A.AGG_MAX, 1)],
                 where=(">", ("col", 1), ("const", 500))),
            dict(group_by=[8], out_cols=[8, 6, 0], aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, 1), (A.AGG_MAX, 7)]),
            dict(group_by=[6], out_cols=[6], aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_AVG, 3)],
                 where=("like", ("col", 6), ("const", "cust1%"))),
            dict(group_by=[4], out_cols=[4, 5], aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_MIN, 0)]),
        ]
        for spec in specs:
            got = pj.join_aggregate(lib_g, og, cg, pc.build(spec, join=(cg, 4, 0)), world=world)
            want = oo.execute(pc.build(spec, join=(co, 4, 0)))
            got.pop("stats")
            pc.compare_results(got, want)
        # the lists partition the rows: every row on exactly one owner's list
        rows, counts, classes = pj.partition(lib_g, og, 4, world)
        assert sum(counts) == og.row_count() and len(set(rows.tolist())) == sum(counts)
        assert classes == 0b011  # NULL keys and numbers
        # an unresolved key column never matches (evaluator_joins.c:54)
        got = pj.join_aggregate(lib_g, og, cg, pc.build(specs[0], join=(cg, -1, 0)), world=world)
        assert got["groups"][0]["count"] == 0


def test_join_keys_of_different_classes_are_declined():
    """value_compare is 0 ("equal") across type classes (src/csv_reader.c:98-130): a numeric key column joined to
    a text one matches every pair in the reference. The hash join cannot reproduce that and must say so, also
    when each side on its own is of one class."""
    left = b"k,v\n1,a\n2,b\n"
    right = b"k,w\nx,1\ny,2\n"
    with Table.from_bytes(left, lib=gpu()) as lg, Table.from_bytes(right, lib=gpu()) as rg:
        with pytest.raises(CqError) as ei:
            lg.execute(Plan(aggs=[(A.AGG_COUNT_STAR, -1)], join=(rg, 0, 0)))
        assert ei.value.code == A.ERR_UNSUPPORTED


def test_partitioned_join_declines_a_stray_key_of_another_class():
    """Left keys [1, 'x'], right keys [1]: the single-GPU join declines (the reference matches 'x' with 1 because
    value_compare is 0 across classes). Hash-partitioned, the text key may land on a rank that owns no other key,
    where no single rank sees the mix: the class masks of all ranks and both sides are ORed and every rank declines."""
    from cq_b200 import partitioned_join as pj
    left = b"k,v\n1,a\nx,b\n2,c\n3,d\n"
    right = b"k,w\n1,10\n2,20\n"
    spec = dict(aggs=[(A.AGG_COUNT_STAR, -1)])
    with Table.from_bytes(left, lib=gpu()) as lg, Table.from_bytes(right, lib=gpu()) as rg:
        with pytest.raises(CqError) as e1:
            lg.execute(pc.build(spec, join=(rg, 0, 0)))
        assert e1.value.code == A.ERR_UNSUPPORTED
        for world in (2, 3, 5):
            with pytest.raises(CqError) as e2:
                pj.join_aggregate(gpu(), lg, rg, pc.build(spec, join=(rg, 0, 0)), world=world)
            assert e2.value.code == A.ERR_UNSUPPORTED


def test_ahead_of_time_lean_kernels_without_the_run_time_compiler():
    """CQG_JIT=0: the same plans on the ahead-of-time (generic) lean kernels, in a fresh process (the switch is
    read once). Every other test of this file runs the kernels compiled per query shape."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, CQG_JIT="0")
    sel = "test_plan_parity and (group_name or filter_not or scalar_aggs or count_age or lean_group or group_high_card)"
    r = subprocess.run([sys.executable, "-m", "pytest", __file__, "-q", "-m", "gpu", "-x", "-k", sel], env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout
