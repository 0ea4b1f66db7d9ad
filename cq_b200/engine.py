"""Thin Python host layer over the C-ABI (include/cq_gpu.h): tables, query plans, results.

This is plumbing for tests and bench.py; the drop-in seam for cq itself is the C dispatcher
cq_b200/host/cq_dispatch.c, which builds the same cqg_query_t from cq's AST.

The product backend is libcqgpu.so (CUDA, sm_100a). There is no CPU route in this module:
`gpu()` raises when the library is missing or no device is visible.
"""
import ctypes as C
import os

from . import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcqgpu.so")
_gpu = None


class CqError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[cqg error {code}] {msg}")
        self.code = code


def gpu():
    """The CUDA backend. Fails loudly when libcqgpu.so has not been built."""
    global _gpu
    if _gpu is None:
        if not os.path.exists(LIB_PATH):
            raise CqError(A.ERR_CUDA, f"{LIB_PATH} is missing: run `python -m cq_b200.build` (there is no CPU fallback)")
        _gpu = A.Lib(LIB_PATH, "cqg_")
    return _gpu


def _check(lib, rc):
    if rc != A.OK:
        raise CqError(rc, (lib.last_error() or b"").decode("utf-8", "replace"))


def csv_config(delimiter=",", quote='"', has_header=True):
    c = A.CsvConfig()
    c.delimiter = delimiter.encode("latin1")
    c.quote = quote.encode("latin1")
    c.has_header = b"\x01" if has_header else b"\x00"
    c.reserved = b"\x00"
    return c


# ---------------------------------------------------------------------------------------
# values
# ---------------------------------------------------------------------------------------
def py_value(v):
    """cqg_value_t -> ('N',) | ('I', int) | ('D', float) | ('S', bytes) | ('T', y, m, d)"""
    t = v.type
    if t == A.TYPE_INTEGER:
        return ("I", v.int_value)
    if t == A.TYPE_DOUBLE:
        return ("D", v.double_value)
    if t == A.TYPE_STRING:
        return ("S", v.string_value if v.string_value is not None else b"")
    if t == A.TYPE_DATE:
        return ("T", v.date_value.year, v.date_value.month, v.date_value.day)
    return ("N",)


def make_value(x, keep):
    """Python literal -> cqg_value_t. None, int, float, bytes/str, ('T', y, m, d)."""
    v = A.Value()
    if x is None:
        v.type = A.TYPE_NULL
    elif isinstance(x, bool):
        raise TypeError("bool constant")
    elif isinstance(x, int):
        v.type = A.TYPE_INTEGER
        v.int_value = x
    elif isinstance(x, float):
        v.type = A.TYPE_DOUBLE
        v.double_value = x
    elif isinstance(x, (bytes, str)):
        b = x.encode("utf-8") if isinstance(x, str) else x
        buf = C.create_string_buffer(b)
        keep.append(buf)
        v.type = A.TYPE_STRING
        v.string_value = C.cast(buf, C.c_char_p)
    elif isinstance(x, tuple) and x and x[0] == "T":
        v.type = A.TYPE_DATE
        v.date_value.year, v.date_value.month, v.date_value.day = x[1], x[2], x[3]
    else:
        raise TypeError(f"constant {x!r}")
    return v


# ---------------------------------------------------------------------------------------
# predicate builder: nested tuples -> postfix cqg_insn_t
#   ("col", i) ("const", literal) ("+", a, b) ... ("neg", a)
#   ("=", a, b) ("!=",..) (">",..) ("<",..) (">=",..) ("<=",..) ("like", a, b) ("ilike", a, b)
#   ("in", a, [items]) ("not in", a, [items]) ("and", p, q) ("or", p, q) ("not", p) ("true",) ("false",)
# ---------------------------------------------------------------------------------------
_ARITH = {"+": A.OP_ADD, "-": A.OP_SUB, "*": A.OP_MUL, "/": A.OP_DIV, "%": A.OP_MOD, "&": A.OP_BAND, "|": A.OP_BOR,
          "^": A.OP_BXOR, "?": A.OP_ARITH_NULL}
_CMP = {"=": A.OP_EQ, "!=": A.OP_NE, ">": A.OP_GT, "<": A.OP_LT, ">=": A.OP_GE, "<=": A.OP_LE, "like": A.OP_LIKE,
        "ilike": A.OP_ILIKE}


def _emit(node, code, consts):
    k = node[0]
    if k == "col":
        code.append((A.OP_COL, node[1]))
    elif k == "const":
        consts.append(node[1])
        code.append((A.OP_CONST, len(consts) - 1))
    elif k in _ARITH and len(node) == 3:
        _emit(node[1], code, consts)
        _emit(node[2], code, consts)
        code.append((_ARITH[k], 0))
    elif k == "neg":
        _emit(node[1], code, consts)
        code.append((A.OP_NEG, 0))
    elif k in _CMP:
        _emit(node[1], code, consts)
        _emit(node[2], code, consts)
        code.append((_CMP[k], 0))
    elif k in ("in", "not in"):
        _emit(node[1], code, consts)
        for it in node[2]:
            _emit(it, code, consts)
        code.append((A.OP_IN if k == "in" else A.OP_NOT_IN, len(node[2])))
    elif k in ("and", "or"):
        _emit(node[1], code, consts)
        _emit(node[2], code, consts)
        code.append((A.OP_AND if k == "and" else A.OP_OR, 0))
    elif k == "not":
        _emit(node[1], code, consts)
        code.append((A.OP_NOT, 0))
    elif k == "true":
        code.append((A.OP_TRUE, 0))
    elif k == "false":
        code.append((A.OP_FALSE, 0))
    else:
        raise ValueError(f"predicate node {node!r}")


class Plan:
    """Owns a cqg_query_t and everything it points to."""

    def __init__(self, where=None, group_by=(), aggs=(), out_cols=(), mode="aggregate", max_rows=-1, join=None):
        self._keep = []
        q = A.Query()
        q.mode = A.MODE_AGGREGATE if mode == "aggregate" else A.MODE_SELECT
        if where is not None:
            code, consts = [], []
            _emit(where, code, consts)
            carr = (A.Insn * len(code))(*[A.Insn(op, a) for op, a in code])
            varr = (A.Value * max(len(consts), 1))(*[make_value(c, self._keep) for c in consts])
            self._keep += [carr, varr]
            q.where.code = C.cast(carr, C.POINTER(A.Insn))
            q.where.n_code = len(code)
            q.where.consts = C.cast(varr, C.POINTER(A.Value))
            q.where.n_consts = len(consts)
        q.n_group_cols = len(group_by)
        for i, c in enumerate(group_by):
            q.group_cols[i] = c
        q.n_aggs = len(aggs)
        for i, (f, c) in enumerate(aggs):
            q.aggs[i].func = f
            q.aggs[i].col = c
        q.n_out_cols = len(out_cols)
        for i, c in enumerate(out_cols):
            q.out_cols[i] = c
        q.max_rows = max_rows
        if join is not None:
            right, lcol, rcol = join[:3]
            self._keep.append(right)
            q.join.right = right.handle
            q.join.left_col = lcol
            q.join.right_col = rcol
            q.join.type = join[3] if len(join) > 3 else A.JOIN_INNER
        self.q = q


class Table:
    def __init__(self, lib, handle, keep=None):
        self.lib = lib
        self.handle = handle
        self._keep = keep

    @classmethod
    def open(cls, path, cfg=None, lib=None):
        lib = lib or gpu()
        h = C.c_void_p()
        _check(lib, lib.table_open(os.fsencode(path), cfg or csv_config(), C.byref(h)))
        return cls(lib, h)

    @classmethod
    def from_bytes(cls, data, cfg=None, lib=None, pinned=False):
        """`data`: bytes, or (address, size) of a host buffer (pinned=True when page-locked)."""
        lib = lib or gpu()
        h = C.c_void_p()
        if isinstance(data, tuple):
            addr, size = data
            keep = None
        else:
            keep = C.create_string_buffer(data, len(data))
            addr, size = C.addressof(keep), len(data)
        _check(lib, lib.table_open_buffer(C.c_void_p(addr), size, 1 if pinned else 0, cfg or csv_config(), C.byref(h)))
        return cls(lib, h, keep)

    @classmethod
    def from_device(cls, ptr, size, cfg=None, lib=None, keep=None):
        lib = lib or gpu()
        h = C.c_void_p()
        _check(lib, lib.table_open_device(ptr, size, cfg or csv_config(), C.byref(h)))
        return cls(lib, h, keep)

    def close(self):
        if self.handle:
            self.lib.table_close(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_shard(self, index, count):
        _check(self.lib, self.lib.table_set_shard(self.handle, index, count))

    def set_global_offset(self, off):
        _check(self.lib, self.lib.table_set_global_offset(self.handle, off))

    def set_gpus(self, n):
        """Spread the table over the first `n` devices of this process (cqg_table_set_gpus); before the first query."""
        _check(self.lib, self.lib.table_set_gpus(self.handle, n))
        return self.lib.table_gpus(self.handle)

    @property
    def columns(self):
        n = self.lib.table_column_count(self.handle)
        return [self.lib.table_column_name(self.handle, i).decode("utf-8", "replace") for i in range(n)]

    def column_index(self, name):
        return self.lib.table_column_index(self.handle, name.encode())

    @property
    def size(self):
        return self.lib.table_size(self.handle)

    def row_count(self):
        n = C.c_int64()
        _check(self.lib, self.lib.table_row_count(self.handle, C.byref(n)))
        return n.value

    def execute(self, plan):
        r = C.POINTER(A.Result)()
        _check(self.lib, self.lib.execute(self.handle, C.byref(plan.q), C.byref(r)))
        try:
            return decode_result(r.contents, plan)
        finally:
            self.lib.result_free(r)

    def execute_raw(self, plan):
        """Run and return only the scalar facts (for timing loops)."""
        r = C.POINTER(A.Result)()
        _check(self.lib, self.lib.execute(self.handle, C.byref(plan.q), C.byref(r)))
        c = r.contents
        out = {"n_groups": c.n_groups, "n_selected": c.n_selected, "rows_scanned": c.rows_scanned,
               "kernel_ms": c.kernel_ms, "kernel_launches": c.kernel_launches,
               "count0": c.count[0] if c.n_groups > 0 else 0}
        self.lib.result_free(r)
        return out


def decode_result(c, plan):
    q = plan.q
    if q.mode == A.MODE_AGGREGATE:
        G = c.n_groups
        groups = []
        for g in range(G):
            groups.append({
                "first_offset": c.first_offset[g],
                "count": c.count[g],
                "aggs": [py_value(c.value[a * G + g]) for a in range(q.n_aggs)],
                "sum": [c.sum[a * G + g] for a in range(q.n_aggs)],
                "ncount": [c.ncount[a * G + g] for a in range(q.n_aggs)],
                "out": [py_value(c.out[k * G + g]) for k in range(q.n_out_cols)],
            })
        return {"groups": groups, "rows_scanned": c.rows_scanned, "kernel_ms": c.kernel_ms,
                "kernel_launches": c.kernel_launches}
    n = c.n_rows_out
    rows = [[py_value(c.rows[i * q.n_out_cols + k]) for k in range(q.n_out_cols)] for i in range(n)]
    return {"n_selected": c.n_selected, "rows": rows, "row_offset": [c.row_offset[i] for i in range(n)],
            "row_offset_right": [c.row_offset_right[i] for i in range(n)] if c.row_offset_right else None,
            "rows_scanned": c.rows_scanned, "kernel_ms": c.kernel_ms, "kernel_launches": c.kernel_launches}
