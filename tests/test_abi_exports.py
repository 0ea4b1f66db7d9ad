"""CPU-side checks of the boundary: libcqgpu.so loads and exports every symbol that
include/cq_gpu.h declares; struct layouts match the header; the product has no CPU route."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from cq_b200 import _abi as A
from cq_b200 import engine

HEADER = os.path.join(ROOT, "include", "cq_gpu.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cqg_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported():
    assert os.path.exists(engine.LIB_PATH), "libcqgpu.so not built: python -m cq_b200.build"
    dll = C.CDLL(engine.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(dll, n), f"{n} declared in include/cq_gpu.h but not exported"


def test_python_prototypes_cover_header():
    have = {"cqg_" + k for k in A.PROTOTYPES}
    assert set(declared_symbols()) <= have


def test_struct_layouts():
    assert C.sizeof(A.Value) == 24
    assert C.sizeof(A.Insn) == 8
    assert C.sizeof(A.CsvConfig) == 4
    assert A.Query.group_cols.offset % 4 == 0
    assert C.sizeof(A.Agg) == 8


def test_no_device_fails_loudly():
    """Without a CUDA device the library must fail with CQG_ERR_CUDA, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = engine.gpu()
    with pytest.raises(engine.CqError) as ei:
        engine.Table.from_bytes(b"a,b\n1,2\n", lib=lib)
    assert ei.value.code == A.ERR_CUDA


def test_product_does_not_reference_oracle():
    """Nothing under cq_b200/ may import, link or call oracle/ (cq_dispatch.c's cqo_ binding is
    compiled only with -DCQ_BACKEND_ORACLE, by oracle/Makefile)."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "cq_b200")):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".c", ".h")) and f != "Makefile":
                continue
            text = open(os.path.join(dirpath, f), errors="replace").read()
            if f.endswith((".c", ".cu", ".cuh", ".h")):
                text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            if f == "cq_dispatch.c":
                text = re.sub(r"#ifdef CQ_BACKEND_ORACLE.*?#else", "", text, flags=re.S)
            if f in ("build.py", "cq_dump.c"):
                continue
            if re.search(r"liboracle|cqo_|oracle/", text):
                bad.append(f)
    assert not bad, bad
    out = os.popen(f"nm -D {engine.LIB_PATH} | grep -c cqo_").read().strip()
    assert out == "0"
