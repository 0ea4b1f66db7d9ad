"""Test-side loader of the CPU restatement (oracle/liboracle.so, prefix cqo_). Checker only."""
import ctypes as C
import os
import subprocess

from cq_b200 import _abi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
_ORACLE_FUNCS = {"last_error", "table_open", "table_open_buffer", "table_set_shard", "table_close", "table_column_count",
                 "table_column_name", "table_column_index", "table_size", "execute", "result_free", "table_row_count",
                 "parse_value", "value_release", "generate_bigdata_bound"}
_lib = None


def oracle():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
        _lib = A.Lib(ORACLE_SO, "cqo_", only=_ORACLE_FUNCS)
        g = _lib.dll.cqo_generate_bigdata
        g.restype = C.c_int
        g.argtypes = [C.c_void_p, C.c_size_t, C.c_int64, C.c_uint64, C.c_int64, C.POINTER(C.c_size_t)]
        _lib.generate_bigdata_host = g
    return _lib


def generate_bigdata(rows, seed=1, key_card=0):
    """bytes of the seeded restatement of utils/generate_big_dataset.py (CPU generator)."""
    lib = oracle()
    cap = lib.generate_bigdata_bound(rows, key_card)
    buf = C.create_string_buffer(cap)
    n = C.c_size_t()
    rc = lib.generate_bigdata_host(buf, cap, rows, seed, key_card, C.byref(n))
    assert rc == 0
    return buf.raw[:n.value]
