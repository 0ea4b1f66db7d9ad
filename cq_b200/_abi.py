"""ctypes view of include/cq_gpu.h (the C-ABI of libcqgpu.so).

Only declarations live here: struct layouts, enum values and function prototypes; `Lib`
binds them to a shared library under a symbol prefix (``cqg_`` for the product).
"""
import ctypes as C

# ---- enums (include/cq_gpu.h) ----
TYPE_NULL, TYPE_INTEGER, TYPE_DOUBLE, TYPE_STRING, TYPE_DATE = 0, 1, 2, 3, 4
OK, ERR_CUDA, ERR_IO, ERR_ARG, ERR_UNSUPPORTED, ERR_NOMEM, ERR_UNSUPPORTED_PLAN = 0, 1, 2, 3, 4, 5, 6

OP_COL, OP_CONST = 1, 2
OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_MOD, OP_BAND, OP_BOR, OP_BXOR, OP_NEG, OP_POS, OP_ARITH_NULL = range(10, 21)
OP_EQ, OP_NE, OP_GT, OP_LT, OP_GE, OP_LE, OP_IN, OP_NOT_IN, OP_LIKE, OP_ILIKE = range(30, 40)
OP_AND, OP_OR, OP_NOT, OP_TRUE, OP_FALSE, OP_POP = range(50, 56)

AGG_COUNT_STAR, AGG_COUNT, AGG_SUM, AGG_AVG, AGG_MIN, AGG_MAX = range(6)
JOIN_INNER, JOIN_LEFT, JOIN_RIGHT, JOIN_FULL = 0, 1, 2, 3
MODE_AGGREGATE, MODE_SELECT = 0, 1
MAX_GROUP_COLS, MAX_AGGS, MAX_OUT_COLS = 8, 16, 64


class Date(C.Structure):
    _fields_ = [("year", C.c_int), ("month", C.c_int), ("day", C.c_int)]


class _ValueU(C.Union):
    _fields_ = [("int_value", C.c_longlong), ("double_value", C.c_double), ("string_value", C.c_char_p),
                ("date_value", Date)]


class Value(C.Structure):
    _anonymous_ = ("u",)
    _fields_ = [("type", C.c_int32), ("reserved", C.c_int32), ("u", _ValueU)]


assert C.sizeof(Value) == 24


class CsvConfig(C.Structure):
    _fields_ = [("delimiter", C.c_char), ("quote", C.c_char), ("has_header", C.c_char), ("reserved", C.c_char)]


class Insn(C.Structure):
    _fields_ = [("op", C.c_int32), ("a", C.c_int32)]


class Predicate(C.Structure):
    _fields_ = [("code", C.POINTER(Insn)), ("n_code", C.c_int32), ("consts", C.POINTER(Value)), ("n_consts", C.c_int32)]


class Agg(C.Structure):
    _fields_ = [("func", C.c_int32), ("col", C.c_int32)]


class Join(C.Structure):
    _fields_ = [("right", C.c_void_p), ("left_col", C.c_int32), ("right_col", C.c_int32), ("type", C.c_int32),
                ("reserved", C.c_int32)]


class Query(C.Structure):
    _fields_ = [
        ("mode", C.c_int32),
        ("where", Predicate),
        ("join", Join),
        ("n_group_cols", C.c_int32),
        ("group_cols", C.c_int32 * MAX_GROUP_COLS),
        ("n_aggs", C.c_int32),
        ("aggs", Agg * MAX_AGGS),
        ("n_out_cols", C.c_int32),
        ("out_cols", C.c_int32 * MAX_OUT_COLS),
        ("max_rows", C.c_int64),
    ]


class Result(C.Structure):
    _fields_ = [
        ("n_groups", C.c_int64),
        ("first_offset", C.POINTER(C.c_uint64)),
        ("count", C.POINTER(C.c_int64)),
        ("sum", C.POINTER(C.c_double)),
        ("ncount", C.POINTER(C.c_int64)),
        ("value", C.POINTER(Value)),
        ("out", C.POINTER(Value)),
        ("n_selected", C.c_int64),
        ("n_rows_out", C.c_int64),
        ("row_offset", C.POINTER(C.c_uint64)),
        ("row_offset_right", C.POINTER(C.c_uint64)),
        ("rows", C.POINTER(Value)),
        ("rows_scanned", C.c_int64),
        ("n_aggs", C.c_int32),
        ("n_out_cols", C.c_int32),
        ("kernel_ms", C.c_double),
        ("kernel_launches", C.c_int32),
        ("arena", C.c_void_p),
    ]


# name -> (restype, argtypes); every symbol include/cq_gpu.h declares
PROTOTYPES = {
    "last_error": (C.c_char_p, []),
    "device_count": (C.c_int, []),
    "set_device": (C.c_int, [C.c_int]),
    "abi_version": (C.c_int, []),
    "table_open": (C.c_int, [C.c_char_p, CsvConfig, C.POINTER(C.c_void_p)]),
    "table_open_buffer": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, CsvConfig, C.POINTER(C.c_void_p)]),
    "table_open_device": (C.c_int, [C.c_uint64, C.c_size_t, CsvConfig, C.POINTER(C.c_void_p)]),
    "device_padding": (C.c_size_t, []),
    "table_set_shard": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "table_set_global_offset": (C.c_int, [C.c_void_p, C.c_uint64]),
    "table_close": (None, [C.c_void_p]),
    "table_column_count": (C.c_int, [C.c_void_p]),
    "table_column_name": (C.c_char_p, [C.c_void_p, C.c_int]),
    "table_column_index": (C.c_int, [C.c_void_p, C.c_char_p]),
    "table_size": (C.c_size_t, [C.c_void_p]),
    "table_device_ptr": (C.c_uint64, [C.c_void_p]),
    "execute": (C.c_int, [C.c_void_p, C.POINTER(Query), C.POINTER(C.POINTER(Result))]),
    "result_free": (None, [C.POINTER(Result)]),
    "table_row_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "parse_value": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(Value)]),
    "value_release": (None, [C.POINTER(Value)]),
    "execute_partial": (C.c_int, [C.c_void_p, C.POINTER(Query), C.POINTER(C.c_void_p)]),
    "partial_kernel_ms": (C.c_double, [C.c_void_p]),
    "partial_rows_scanned": (C.c_int64, [C.c_void_p]),
    "partial_record_size": (C.c_size_t, [C.c_void_p]),
    "partial_count": (C.c_int64, [C.c_void_p]),
    "partial_export": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_int64, C.POINTER(C.c_int64)]),
    "partial_owner_counts": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]),
    "partial_new_like": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "partial_merge": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int64]),
    "partial_finish": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.POINTER(Result))]),
    "partial_free": (None, [C.c_void_p]),
    "partition_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "rowlist_device_ptr": (C.c_uint64, [C.c_void_p]),
    "rowlist_key_classes": (C.c_uint, [C.c_void_p]),
    "rowlist_counts": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]),
    "rowlist_copy": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int64]),
    "rowlist_free": (None, [C.c_void_p]),
    "execute_partial_rows": (C.c_int, [C.c_void_p, C.POINTER(Query), C.c_uint64, C.c_int64, C.c_uint64, C.c_int64,
                                       C.POINTER(C.c_void_p)]),
    "generate_bigdata": (C.c_int, [C.c_uint64, C.c_size_t, C.c_int64, C.c_uint64, C.c_int64, C.POINTER(C.c_size_t)]),
    "generate_bigdata_range": (C.c_int, [C.c_uint64, C.c_size_t, C.c_int64, C.c_int64, C.c_uint64, C.c_int64, C.c_int,
                                         C.POINTER(C.c_size_t)]),
    "generate_bigdata_bound": (C.c_size_t, [C.c_int64, C.c_int64]),
    "table_set_gpus": (C.c_int, [C.c_void_p, C.c_int]),
    "table_gpus": (C.c_int, [C.c_void_p]),
    "table_field_counts": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.c_int64, C.POINTER(C.c_int32)]),
    "total_kernel_launches": (C.c_int64, []),
    "kernel_launches_named": (C.c_int64, [C.c_char_p]),
    "last_scan_stats": (None, [C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
}


class Lib:
    """A loaded backend: ``Lib(path, 'cqg_')`` for libcqgpu.so."""

    def __init__(self, path, prefix, only=None):
        self.path = path
        self.prefix = prefix
        self.dll = C.CDLL(path)
        for name, (res, args) in PROTOTYPES.items():
            if only is not None and name not in only:
                continue
            fn = getattr(self.dll, prefix + name)
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)
