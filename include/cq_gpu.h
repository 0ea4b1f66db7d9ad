/*
 * cq_gpu.h — C-ABI of libcqgpu: the B200-native replacement for cq's data-parallel
 * query hot path (CSV bytes -> row/field boundaries -> typed decode -> WHERE ->
 * GROUP BY aggregates / equi-JOIN).
 *
 * Plain C: pointers, sizes, POD structs. No torch / C++ types cross this boundary.
 * Every entry point names the reference interface it replaces (paths relative to
 * the krow89/cq tree).
 *
 * Semantics are the reference's. In particular:
 *   - values are typed PER VALUE (src/csv_reader.c:133-240), not per column;
 *   - value_compare's cross-type "equal" (src/csv_reader.c:98-130) is honoured;
 *   - group keys are the reference's string renderings (evaluator_aggregates.c:121-141);
 *   - groups come back in first-appearance order among the filtered rows.
 *
 * There is no CPU fallback behind these calls: when no CUDA device / kernel image
 * is available they fail with CQG_ERR_CUDA and a message in cqg_last_error().
 */
#ifndef CQ_GPU_H
#define CQ_GPU_H

#ifndef __CUDACC_RTC__ /* (kernels compiled at run time get these types from cqg_rtc.h) */
#include <stddef.h>
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- values: layout-identical to the reference's `Value` (include/csv_reader.h:8-33) ---- */

typedef enum {
    CQG_TYPE_NULL = 0,     /* VALUE_TYPE_NULL    */
    CQG_TYPE_INTEGER = 1,  /* VALUE_TYPE_INTEGER */
    CQG_TYPE_DOUBLE = 2,   /* VALUE_TYPE_DOUBLE  */
    CQG_TYPE_STRING = 3,   /* VALUE_TYPE_STRING  */
    CQG_TYPE_DATE = 4      /* VALUE_TYPE_DATE    */
} cqg_type_t;

typedef struct {
    int year, month, day;
} cqg_date_t;

typedef struct {
    int32_t type; /* cqg_type_t */
    int32_t reserved;
    union {
        long long int_value;
        double double_value;
        char* string_value; /* NUL-terminated, owned by the cqg_result_t / caller */
        cqg_date_t date_value;
    };
} cqg_value_t; /* 24 bytes, same as sizeof(Value) */

/* ---- CSV dialect: the reference's CsvConfig (include/csv_reader.h:66-70) ---- */

typedef struct {
    char delimiter;
    char quote;
    char has_header; /* bool */
    char reserved;
} cqg_csv_config_t;

/* ---- errors: reference convention is "one stderr line + NULL"; here: code + message ---- */

enum {
    CQG_OK = 0,
    CQG_ERR_CUDA = 1,        /* no device, kernel image missing, CUDA runtime error */
    CQG_ERR_IO = 2,          /* open/fstat/mmap failed (src/mmap.c:78-108 returns NULL) */
    CQG_ERR_ARG = 3,         /* malformed plan */
    CQG_ERR_UNSUPPORTED = 4, /* input the kernels met and cannot reproduce exactly (data dependent) */
    CQG_ERR_NOMEM = 5,
    CQG_ERR_UNSUPPORTED_PLAN = 6 /* shape or dialect outside the GPU path, found at plan time before any table byte
                                    was uploaded: the caller keeps the reference route, like for any other shape
                                    that is not an operator of this path */
};

const char* cqg_last_error(void);

/* ---- device / library ---- */

/* number of visible CUDA devices; <0 on error. */
int cqg_device_count(void);
/* select the device used by subsequent calls made from this thread (one process per GPU
 * under torchrun: call with LOCAL_RANK). */
int cqg_set_device(int device);
/* ABI version of this header. */
int cqg_abi_version(void);

/* ---- tables: replaces csv_load / csv_free (src/csv_reader.c:375-490) + portable_mmap ---- */

typedef struct cqg_table cqg_table_t;

/* mmap `path` (src/mmap.c:78) and split the header line (src/csv_reader.c:341-357). The bytes are staged into HBM
 * when a query first needs them (parallel chunked async copies through page-locked bounce buffers), so a statement
 * whose plan is declined never pays for the upload. A dialect the kernels do not cover fails with
 * CQG_ERR_UNSUPPORTED_PLAN. */
int cqg_table_open(const char* path, cqg_csv_config_t cfg, cqg_table_t** out);

/* same, but the CSV bytes are already in host memory (`data` must stay valid until close).
 * If `pinned` is non-zero the buffer is page-locked and is DMA'd directly. */
int cqg_table_open_buffer(const void* data, size_t size, int pinned, cqg_csv_config_t cfg,
                          cqg_table_t** out);

/* the CSV bytes are already resident in HBM at `device_ptr` (no host copy is made; the
 * allocation must have >= cqg_device_padding() writable bytes after `size`). */
int cqg_table_open_device(uint64_t device_ptr, size_t size, cqg_csv_config_t cfg,
                          cqg_table_t** out);
size_t cqg_device_padding(void);

/* Multi-GPU residency in ONE process (the drop-in CLI: `CQ_GPUS=8 cq -q ...`; cqg_table_open reads CQ_GPUS itself). The
 * file gets one virtual address range whose pages live on the first `ngpu` devices (slice d on device d, CUDA virtual
 * memory management, every device mapped read/write); staging runs one copy thread per device. cqg_execute then runs
 * aggregates without a join on all devices (each scans the rows starting in its slice, the partial group records cross
 * NVLink to device 0, merge and finish there) and any other shape on device 0 over the whole range. Call before the
 * first query; tables below CQG_MULTI_MIN_BYTES (64 MB) stay on one device. Replaces nothing in the reference (it has no
 * threads): src/mmap.c:78-108 + src/csv_reader.c:375-465 are the single-core load it stands in for. */
int cqg_table_set_gpus(cqg_table_t* t, int ngpu);
/* devices the table is (or will be) resident on */
int cqg_table_gpus(const cqg_table_t* t);

/* restrict the table to the rows whose FIRST byte lies in the byte range of shard
 * `index` of `count` equal ranges (exact because row boundaries are not
 * quote-aware, src/csv_reader.c:407). Default is shard 0 of 1. The header is always
 * taken from the start of the file. */
int cqg_table_set_shard(cqg_table_t* t, int index, int count);

/* multi-GPU runs in which every rank holds only its own slice of the file: `offset` is the
 * position of this table's byte 0 in the whole file, added to row offsets so that
 * first-appearance order (evaluator_aggregates.c:159-163) is global. Default 0. */
int cqg_table_set_global_offset(cqg_table_t* t, uint64_t offset);

void cqg_table_close(cqg_table_t* t);

/* table->column_count / columns[i].name (src/csv_reader.c:343-357: trimmed, "$i" when
 * blank or has_header is false). */
int cqg_table_column_count(const cqg_table_t* t);
const char* cqg_table_column_name(const cqg_table_t* t, int col);
/* csv_get_column_index (src/csv_reader.c:500-509): case-insensitive first match, -1. */
int cqg_table_column_index(const cqg_table_t* t, const char* name);
size_t cqg_table_size(const cqg_table_t* t);
uint64_t cqg_table_device_ptr(const cqg_table_t* t);

/* ---- predicate programs: replaces evaluate_condition / evaluate_expression ----
 * (src/evaluator/evaluator_conditions.c:62-164, evaluator_expressions.c:23-42,101-263).
 * Postfix code over a small stack of tagged values. Column operands name a CSV column
 * index of table 0 (or, for joined rows, left columns then right columns, the layout
 * perform_join builds at evaluator_joins.c:73-77). */

typedef enum {
    CQG_OP_COL = 1,    /* push column `a` (a<0: unknown column -> NULL, expressions.c:33-42) */
    CQG_OP_CONST = 2,  /* push consts[a] (literal folded once with parse_value, Q8) */
    CQG_OP_ADD = 10,
    CQG_OP_SUB = 11,
    CQG_OP_MUL = 12,
    CQG_OP_DIV = 13,
    CQG_OP_MOD = 14,
    CQG_OP_BAND = 15,
    CQG_OP_BOR = 16,
    CQG_OP_BXOR = 17,
    CQG_OP_NEG = 18,   /* unary minus (expressions.c:112-122) */
    CQG_OP_POS = 19,   /* unary plus: operand unchanged (expressions.c:123-126) */
    CQG_OP_ARITH_NULL = 20, /* unknown arithmetic operator with numeric operands: result 0.0
                               typed by the int rule (expressions.c:187-260 falls through) */
    CQG_OP_EQ = 30,
    CQG_OP_NE = 31,
    CQG_OP_GT = 32,
    CQG_OP_LT = 33,
    CQG_OP_GE = 34,
    CQG_OP_LE = 35,
    CQG_OP_IN = 36,     /* a = item count; stack: value item1..itemN -> bool */
    CQG_OP_NOT_IN = 37,
    CQG_OP_LIKE = 38,   /* stack: string pattern -> bool (conditions.c:152-161) */
    CQG_OP_ILIKE = 39,
    CQG_OP_AND = 50,
    CQG_OP_OR = 51,
    CQG_OP_NOT = 52,
    CQG_OP_TRUE = 53,
    CQG_OP_FALSE = 54,  /* non-condition node / unknown operator (conditions.c:65,163) */
    CQG_OP_POP = 55     /* discard the value on top (operands of an unknown comparison) */
} cqg_opcode_t;

typedef struct {
    int32_t op; /* cqg_opcode_t */
    int32_t a;
} cqg_insn_t;

typedef struct {
    const cqg_insn_t* code; /* NULL / n_code==0: no WHERE, every row passes */
    int32_t n_code;
    const cqg_value_t* consts; /* STRING consts: string_value NUL-terminated */
    int32_t n_consts;
} cqg_predicate_t;

/* ---- aggregates: replaces evaluate_aggregate (evaluator_aggregates.c:263-326) ---- */

typedef enum {
    CQG_AGG_COUNT_STAR = 0,
    CQG_AGG_COUNT = 1, /* == group row count, NULLs NOT skipped (Q11) */
    CQG_AGG_SUM = 2,
    CQG_AGG_AVG = 3,
    CQG_AGG_MIN = 4,
    CQG_AGG_MAX = 5
} cqg_agg_func_t;

typedef struct {
    int32_t func; /* cqg_agg_func_t */
    int32_t col;  /* CSV column index; -1 = column not found -> NULL result (:274-278) */
} cqg_agg_t;

/* ---- join: replaces perform_join for INNER `ident = ident` (evaluator_joins.c:40-140) ---- */

typedef enum {
    CQG_JOIN_INNER = 0,
    CQG_JOIN_LEFT = 1,  /* + every left row without a match, right columns NULL (evaluator_joins.c:128-139) */
    CQG_JOIN_RIGHT = 2, /* + behind all of those, every right row without a match, left columns NULL (:142-171) */
    CQG_JOIN_FULL = 3
} cqg_join_type_t;

typedef struct {
    const cqg_table_t* right; /* NULL: no join */
    int32_t left_col;         /* key column in the left table; -1 = unresolved -> no row matches */
    int32_t right_col;        /* key column in the right table */
    int32_t type;             /* cqg_join_type_t */
    int32_t reserved;
} cqg_join_t;

/* ---- a query over one table (or one inner equi-join) ---- */

typedef enum {
    CQG_MODE_AGGREGATE = 0, /* aggregates, one group `_all_` when n_group_cols==0
                               (src/evaluator.c:232-247) */
    CQG_MODE_SELECT = 1     /* filter_rows + column fetch (evaluator_utils.c:986-1006) */
} cqg_mode_t;

#define CQG_MAX_GROUP_COLS 8
#define CQG_MAX_AGGS 16
#define CQG_MAX_OUT_COLS 64

typedef struct {
    int32_t mode; /* cqg_mode_t */
    cqg_predicate_t where;
    cqg_join_t join;

    /* GROUP BY columns (create_groups, evaluator_aggregates.c:108-176; composite keys
     * src/evaluator.c:112-212). -1 = unknown column: with one key -> zero groups (:116),
     * in a composite key the part renders as "NULL" (:174-176). */
    int32_t n_group_cols;
    int32_t group_cols[CQG_MAX_GROUP_COLS];

    int32_t n_aggs;
    cqg_agg_t aggs[CQG_MAX_AGGS];

    /* AGGREGATE: bare SELECT columns, valued from the group's first row
     * (evaluator_aggregates.c:679-689). SELECT: the columns to fetch for each surviving
     * row. -1 = unknown column -> NULL. */
    int32_t n_out_cols;
    int32_t out_cols[CQG_MAX_OUT_COLS];

    /* SELECT mode: stop materialising after this many rows (<0: all). The count of
     * matching rows is always exact. */
    int64_t max_rows;
} cqg_query_t;

typedef struct {
    /* AGGREGATE mode: n_groups entries, first-appearance order */
    int64_t n_groups;
    uint64_t* first_offset; /* [n_groups] byte offset (in the left file) of the group's first row */
    int64_t* count;         /* [n_groups] rows in group */
    double* sum;            /* [n_aggs][n_groups] SUM over INTEGER/DOUBLE values, in double */
    int64_t* ncount;        /* [n_aggs][n_groups] how many values were numeric */
    cqg_value_t* value;     /* [n_aggs][n_groups] the finished aggregate as the reference types it:
                               COUNT* -> INTEGER, SUM/AVG -> DOUBLE, MIN/MAX -> that value's own
                               type, NULL when the column is unknown / no non-NULL value */
    cqg_value_t* out;       /* [n_out_cols][n_groups] first-row values */

    /* SELECT mode */
    int64_t n_selected;      /* rows passing WHERE (exact, even when truncated by max_rows) */
    int64_t n_rows_out;      /* rows materialised below */
    uint64_t* row_offset;    /* [n_rows_out] byte offset of the row in the left file; outer joins: 2^45 | right
                                offset for a right row without a match (it has no left row) */
    uint64_t* row_offset_right; /* [n_rows_out] byte offset in the right file (joins), else NULL; ~0 for a left
                                   row without a match */
    cqg_value_t* rows;       /* [n_rows_out][n_out_cols] row-major */

    /* bookkeeping */
    int64_t rows_scanned;    /* data rows seen in this shard (after header) */
    int32_t n_aggs, n_out_cols;
    double kernel_ms;        /* device time of the scan kernels (CUDA events) */
    int32_t kernel_launches; /* kernels launched for this query */
    void* arena;             /* owns every pointer above */
} cqg_result_t;

/* Run the query on the GPU. Replaces the interior of evaluate_query_internal between
 * "table loaded" and "ResultSet built" (src/evaluator.c:61-261). */
int cqg_execute(const cqg_table_t* t, const cqg_query_t* q, cqg_result_t** out);
void cqg_result_free(cqg_result_t* r);

/* table->row_count (data rows; src/csv_reader.c:404-427). */
int cqg_table_row_count(const cqg_table_t* t, int64_t* out);

/* parse_value (src/csv_reader.c:195-240) computed by the device decode routine on a
 * one-field launch; STRING results are malloc'ed (free with cqg_value_release). */
int cqg_parse_value(const char* str, size_t len, cqg_value_t* out);
void cqg_value_release(cqg_value_t* v);

/* ---- multi-GPU partial aggregates ----
 * A partial is the shard-local group table serialised to fixed-size records in device
 * memory, so that ranks can exchange it with NCCL (all_gather for few groups, all_to_all
 * by owner = hash % world for many) and fold it back in with cqg_partial_merge. */

typedef struct cqg_partial cqg_partial_t;

/* like cqg_execute(AGGREGATE) but keeps the group table on the device. With a join, the build side is
 * the whole right table on every rank (replicated) and the left table is this rank's shard; the right
 * table, and the partial given to cqg_partial_new_like, must stay open until the merged partial is freed. */
int cqg_execute_partial(const cqg_table_t* t, const cqg_query_t* q, cqg_partial_t** out);
/* device time of the partial's scan kernels (CUDA events) and the data rows it saw */
double cqg_partial_kernel_ms(const cqg_partial_t* p);
int64_t cqg_partial_rows_scanned(const cqg_partial_t* p);
/* record size in bytes and number of records of this partial */
size_t cqg_partial_record_size(const cqg_partial_t* p);
int64_t cqg_partial_count(const cqg_partial_t* p);
/* serialise the records owned by `owner` of `world` (hash % world; world<=1: all) into
 * device memory at dst (capacity in records); returns the number written in *n_out. */
int cqg_partial_export(const cqg_partial_t* p, int owner, int world, uint64_t dst_device_ptr,
                       int64_t capacity, int64_t* n_out);
/* per-owner record counts (host array of `world` int64) for sizing the all-to-all */
int cqg_partial_owner_counts(const cqg_partial_t* p, int world, int64_t* counts);
/* fold `n` serialised records (device memory, any rank's) into a fresh/accumulating table */
int cqg_partial_new_like(const cqg_partial_t* like, cqg_partial_t** out);
int cqg_partial_merge(cqg_partial_t* p, uint64_t src_device_ptr, int64_t n);
/* finish: sort by first offset, decode representative rows, build the result. Offsets in
 * records are GLOBAL file offsets, so `t` must view the whole file (any shard setting). */
int cqg_partial_finish(const cqg_partial_t* p, const cqg_table_t* t, cqg_result_t** out);
void cqg_partial_free(cqg_partial_t* p);

/* ---- hash-partitioned equi-join across ranks (replaces the O(L*R) loop of perform_join,
 * src/evaluator/evaluator_joins.c:96-140, on N GPUs) ----
 * Every rank holds both files whole in HBM (they are read in place: a join moves row REFERENCES, not rows).
 * The work is split by key: rank r scans its byte-range shard of each file and splits the row offsets by
 * owner = hash(canonical join key) % world (cqg_partition_rows); the ranks exchange the lists (NCCL
 * all-to-all of 8-byte global offsets over NVLink); every rank then builds the join table over the RIGHT
 * rows it owns, probes it with the LEFT rows it owns and aggregates (cqg_execute_partial_rows). Equal keys
 * meet on exactly one rank (NULL = NULL and 1 = 1.0 included: both sides canonicalise the key the way
 * value_compare, src/csv_reader.c:98-130, equates them), so the partials merge like any others
 * (cqg_partial_export / _merge / _finish). Row order (first appearance, ties) rests on global offsets. */
typedef struct cqg_rowlist cqg_rowlist_t;
/* offsets of the data rows of this table's shard, split by owner of the key in column `key_col`:
 * `world` dense segments in device memory, segment o = counts[o] offsets (uint64, global file offsets). */
int cqg_partition_rows(const cqg_table_t* t, int key_col, int world, cqg_rowlist_t** out);
uint64_t cqg_rowlist_device_ptr(const cqg_rowlist_t* rl);
/* OR of (1 << comparison class) over the key values of the scanned shard: bit 0 NULL, 1 numeric, 2 text, 3 date
 * (value_compare, src/csv_reader.c:98-130, is 0 - "equal" - across classes). The caller must OR the masks of all
 * ranks and both sides and decline the join when more than one non-NULL class appears: a stray key of another
 * class can land on a rank that owns no other key, where no single rank would see the mix. */
unsigned cqg_rowlist_key_classes(const cqg_rowlist_t* rl);
int cqg_rowlist_counts(const cqg_rowlist_t* rl, int world, int64_t* counts);
/* all segments, in owner order, into caller-owned device memory (capacity in offsets) */
int cqg_rowlist_copy(const cqg_rowlist_t* rl, uint64_t dst_device_ptr, int64_t capacity);
void cqg_rowlist_free(cqg_rowlist_t* rl);
/* cqg_execute_partial for an equi-join query restricted to `n_left` rows of `t` and `n_right` rows of
 * q->join.right (device arrays of file offsets; both tables must view their whole file at offset 0). */
int cqg_execute_partial_rows(const cqg_table_t* t, const cqg_query_t* q, uint64_t left_rows_device_ptr, int64_t n_left,
                             uint64_t right_rows_device_ptr, int64_t n_right, cqg_partial_t** out);

/* ---- csv_load for callers that need the reference's array of structs (DML, the reference's own evaluator) ----
 * Fields per row exactly as parse_line counts them (src/csv_reader.c:278-338) = Row::column_count (:358-366) for the rows
 * at the given byte offsets (host array; cqg_result_t::row_offset of a `SELECT *` projection). Together with that
 * projection it is everything csv_load (src/csv_reader.c:375-465) returns: cq_dispatch.c's csv_load is built from the two. */
int cqg_table_field_counts(const cqg_table_t* t, const uint64_t* row_offsets, int64_t n, int32_t* counts);

/* ---- synthetic data: seeded restatement of utils/generate_big_dataset.py:9-19 ----
 * Fills device memory with header + rows `name,surname,age,gender,height\n`; returns
 * the byte size. key_card>0 appends an integer column `uid` ~ U{0..key_card-1}.
 * Row i depends only on (seed, i), so CPU and GPU generators agree byte for byte. */
int cqg_generate_bigdata(uint64_t device_ptr, size_t capacity, int64_t rows, uint64_t seed,
                         int64_t key_card, size_t* size_out);
/* rows [row_start, row_start + rows) of the same file, with or without the header line: the slice of ONE seeded
 * file a rank holds when the file is split over several GPUs (bench.py --gpus N) */
int cqg_generate_bigdata_range(uint64_t device_ptr, size_t capacity, int64_t row_start, int64_t rows, uint64_t seed,
                               int64_t key_card, int with_header, size_t* size_out);
/* upper bound of the bytes cqg_generate_bigdata needs for `rows` */
size_t cqg_generate_bigdata_bound(int64_t rows, int64_t key_card);

/* ---- introspection for bench.py ---- */
/* total kernels launched by this library in this process */
int64_t cqg_total_kernel_launches(void);
/* launches of one scan kernel family so far: "scan" (general), "lean", "lean2", "lean2g", "lean2k", "leanhc"
 * (cq_b200/csrc/cqg_*.cuh); -1: no such family. Tests use it to assert which kernel a plan ran on. */
int64_t cqg_kernel_launches_named(const char* family);
/* the last scan of this thread that ran on a lean kernel: 16 KB tiles it covered, tiles and single rows it handed over to
 * the general kernel (bytes the lean kernels do not classify: CR, quotes, controls; rows of 64 bytes and more, fields
 * outside their repertoire). bench.py reports the shares. */
void cqg_last_scan_stats(int64_t* tiles, int64_t* handed_tiles, int64_t* handed_rows);

#ifdef __cplusplus
}
#endif
#endif /* CQ_GPU_H */
