"""Drop-in parity: the cq host code (parser, planner, post passes; compiled from the reference
sources) linked with cq_dispatch.c + libcqgpu.so (build/cq_gpu_dump) must print what the
unmodified reference printed (tests/golden/sql_golden.json) for every statement of the corpus.
Counts, keys, integer results, MIN/MAX, row sets and order: bit-exact. DOUBLE cells: 1e-12
relative (SUM/AVG reduction order differs on the GPU)."""
import os

import pytest

from conftest import ROOT, dumps_equal, load_golden, parse_dump, run_dump

pytestmark = pytest.mark.gpu
GPU_DUMP = os.path.join(ROOT, "build", "cq_gpu_dump")
CASES = load_golden()


@pytest.mark.parametrize("case", CASES, ids=[str(c["id"]) for c in CASES])
def test_sql_matches_reference_golden(case):
    if not os.path.exists(GPU_DUMP):
        pytest.skip("build/cq_gpu_dump not built (needs the reference sources at build time)")
    rc, out, err = run_dump(GPU_DUMP, case, env={"CQ_GPU_TRACE": "1"})
    ok, why = dumps_equal(parse_dump(out), parse_dump(case["expected"]), rel=1e-12)
    assert ok, f"{case['sql']}: {why}\n--- got\n{out}\n--- want\n{case['expected']}\n{err}"
    if case.get("route") == "gpu":
        assert "route=gpu" in err, f"{case['sql']} did not take the GPU route:\n{err}"
    if case.get("load") == "gpu":  # a statement on the reference's evaluator: its tables came through the GPU-backed csv_load
        assert "csv_load=gpu" in err and "csv_load=reference" not in err, f"{case['sql']}:\n{err}"
    if case.get("load") == "reference":
        assert "csv_load=reference" in err, f"{case['sql']}:\n{err}"
