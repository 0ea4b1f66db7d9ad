"""Multi-GPU tables in one process (cqg_table_set_gpus / CQ_GPUS): one address range whose pages live on N devices, every
device scanning its own slice, partial records merged on device 0. With fewer than N devices in the box the slices all go
to device 0 (CQG_MULTI_SAME_DEVICE): same code path - ranges, threads, merge, finish through the one range - on one GPU.
Results must equal the oracle's on the whole file, bit for bit (SUM/AVG within 1e-12)."""
import os

import pytest

import parity_cases as pc
from cq_b200 import _abi as A
from cq_b200.engine import Table, gpu
from oracle_lib import generate_bigdata, oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    old = {k: os.environ.get(k) for k in ("CQG_MULTI_MIN_BYTES", "CQG_JIT_MIN_BYTES", "CQG_MULTI_SAME_DEVICE")}
    os.environ["CQG_MULTI_MIN_BYTES"] = "0"
    os.environ["CQG_JIT_MIN_BYTES"] = "0"
    yield torch.cuda.device_count()
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


@pytest.mark.parametrize("ngpu", [2, 3, 8])
def test_multi_gpu_table_equals_the_whole_file(env, ngpu, tmp_path):
    if env < ngpu:
        os.environ["CQG_MULTI_SAME_DEVICE"] = "1"
    else:
        os.environ.pop("CQG_MULTI_SAME_DEVICE", None)
    data = generate_bigdata(400_000, seed=21)  # ~12 MB: six 2 MB pages
    path = tmp_path / "multi.csv"
    path.write_bytes(data)
    lib = gpu()
    names = ["count_age_gt_40", "group_name", "scalar_aggs", "group_gender_minmax", "group_name_surname", "group_high_card",
             "lean_group_two_keys", "scalar_minmax", "select_rows", "select_limit"]
    with Table.open(str(path), lib=lib) as tg, Table.from_bytes(data, lib=oracle()) as to:
        assert tg.set_gpus(ngpu) == ngpu
        assert tg.row_count() == to.row_count() == 400_000
        for name in names:
            spec = pc.plans()[name]
            pc.compare_results(tg.execute(pc.build(spec)), to.execute(pc.build(spec)))


def test_multi_gpu_table_as_a_join_side(env, tmp_path):
    """Shapes that do not shard (joins) run on the table's first device over the one address range."""
    if env < 2:
        os.environ["CQG_MULTI_SAME_DEVICE"] = "1"
    data = generate_bigdata(150_000, seed=22, key_card=5_000)
    right = generate_bigdata(5_000, seed=23, key_card=5_000)
    lp, rp = tmp_path / "l.csv", tmp_path / "r.csv"
    lp.write_bytes(data)
    rp.write_bytes(right)
    UID = 5
    with Table.open(str(lp), lib=gpu()) as tl, Table.open(str(rp), lib=gpu()) as tr, Table.from_bytes(data, lib=oracle()) as ol, \
            Table.from_bytes(right, lib=oracle()) as orr:
        assert tl.set_gpus(2) == 2
        spec = dict(group_by=[3], out_cols=[3], aggs=[(A.AGG_COUNT_STAR, -1), (A.AGG_SUM, 2)])
        got = tl.execute(pc.build(spec, join=(tr, UID, UID)))
        want = ol.execute(pc.build(spec, join=(orr, UID, UID)))
        pc.compare_results(got, want)
