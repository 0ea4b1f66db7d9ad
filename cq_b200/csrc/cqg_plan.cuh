// cqg_plan.cuh — the device-side plan of one query: what the host planner (cqg_api.cu) hands
// to the scan kernel by value. Layout of group-table entries lives here too because the
// host sizes and initialises tables with it.
#pragma once
#include "cqg_rtc.h"
#ifndef __CUDACC_RTC__
#include <cstdint>
#endif

#include "cq_gpu.h"
#include "cqg_device.cuh"

namespace cqg {

constexpr int kMaxSlots = 16;    // distinct CSV columns one side of a query may reference
constexpr int kMaxQueryCols = 256;  // left + right columns addressable by a plan

// extra kernel error flags (continuing cqg_device.cuh)
enum : unsigned {
    KERR_KEY_TAB = 512u,       // composite GROUP BY key part contains '\t' (aliasing in the reference's key string)
    KERR_JOIN_FANOUT = 1024u,  // one left row matches >= 65536 right rows
    KERR_STR_LONG = 2048u,     // MIN/MAX string longer than the packed reference allows
    KERR_OFFSET_RANGE = 4096u, // file offset beyond 2^45
    KERR_ROW_LONG = 8192u,     // internal: row does not fit the tile window (handled, not an error)
    KERR_LEAN_ABORT = 16384u,  // internal: the lean kernel met too many rows it does not cover; rerun on the general kernel
    KERR_LEAN_GROUPS = 32768u  // internal: more groups than a CTA dictionary numbers; rerun the lean kernel on the global table
};
constexpr unsigned kFatalMask = KERR_NUMERIC_RANGE | KERR_SEP_OVERFLOW | KERR_KEY_RANGE | KERR_MINMAX_TIE | KERR_STACK |
                                KERR_JOIN_MIXED | KERR_BIGINT | KERR_KEY_TAB | KERR_JOIN_FANOUT | KERR_STR_LONG |
                                KERR_OFFSET_RANGE;

// ---- fused predicate program (host-compiled from cqg_insn_t when no arithmetic is used) ----
enum : int {
    F_CMP = 1,    // a,b operand refs; n = CQG_OP_EQ..LE
    F_IN = 2,     // a operand; b = first index into refs[]; n = item count
    F_NOT_IN = 3,
    F_LIKE = 4,   // a,b
    F_ILIKE = 5,
    F_AND = 6,
    F_OR = 7,
    F_NOT = 8,
    F_TRUE = 9,
    F_FALSE = 10
};
constexpr int kRefConst = 0x4000;  // ref = kRefConst | const index
constexpr int kRefNull = 0x7fff;   // unknown column
constexpr int kRefSlot = 0x2000;   // fused programs address fields by slot: ref = kRefSlot | slot
constexpr int kMaxFused = 24, kMaxFusedRefs = 32, kMaxFusedConsts = 16;

struct FInsn {
    int16_t op, a, b, n;
};

// ---- aggregate state inside a group entry (every state 16-byte aligned) ----
// SUM/AVG : { int64 sum_i ; double sum_d ; uint64 ncount ; int64 sum_3 }                 32 B
//   INTEGER-typed values add into sum_i (exact), DOUBLE-typed into sum_d, short decimals of the simple
//   route into sum_3 as value*1000 (exact); SUM = sum_i + sum_d + sum_3/1000
// MIN/MAX : { first_nonnull ; date ; num_key ; num_okey ; str ; pad }                    48 B
//   first_nonnull = (okey << 2 | class) of the earliest non-NULL value (class 1 numeric, 2 string, 3 date):
//                   the extreme is taken inside that class only (value_compare is 0 across classes)
//   date          = packed (y<<16|m<<8|d) + 1
//   num_key/okey  = value as an ordered double image + the earliest row holding it (128-bit CAS)
//   str           = table bit 63 | offset << 18 | len   (reference into the resident CSV bytes)
// MIN keeps the smallest image (empty = ~0), MAX the largest (empty = 0); num_okey empty = ~0.
constexpr int kMaxLeanLeaf = 6;
struct LeanLeaf {
    int32_t slot, kind, lop, slen;
    uint32_t A[4];
    long long LB[4];
    uint64_t w0, w1;
};

struct AggSpec {
    int32_t func;
    int32_t col;
    int32_t off;   // byte offset of the state inside the entry (-1: COUNT, no state)
    int32_t slot;  // field slot of `col` (-1: reads as NULL)
};

// entry header
constexpr int kOffHash = 0;    // uint64: 0 empty, bit 63 = being initialised
constexpr int kOffFirst = 8;   // uint64: min okey of the group's rows
constexpr int kOffCount = 16;  // uint64 rows
constexpr int kOffTags = 24;   // uint32 key tags (4 bits per part) ; uint32 pad
constexpr int kOffKeys = 32;   // ngc x { uint64 w0 ; uint64 w1 }
constexpr uint64_t kLockBit = 0x8000000000000000ull;

// key part tags
enum : uint32_t { KT_NULL = 0, KT_INT = 1, KT_DBL_POS = 2, KT_DBL_NEG = 3, KT_STR = 4, KT_STR_HASH = 5, KT_DATE = 6, KT_DBL_BIG = 7 };

// ---- packed group lines of the lean GROUP BY in global mode (cqg_leanhc.cuh, config-3 shape: ~10^6 groups) ----
// One line of 32..192 bytes per group instead of the general entry (32 + 16*ngc + 32 per SUM/AVG + 48 per
// MIN/MAX), laid out in 8-byte words so that what a row must READ comes first and in as few 32-byte chunks
// (one 256-bit load each) as possible:
//   word 0      0 empty, 1 being written, else the group's first okey (+ 2^16, so that it is never 0 or 1); its low
//               16 bits (the join rank, always 0 here) hold the key tags. No hash is stored: the key words decide.
//   words 1..   key part g: (w0, w1) in a wide slot, w0 alone in a narrow one (numbers and NULL only: a row whose
//               part is text there is handed over to the general kernel)
//   then        count; per SUM/AVG an int64 sum of value*1000 (values summed == count: rows with a NULL operand are
//               handed over); then, 16-byte aligned, per MIN/MAX { ordered double image, okey of the earliest row
//               holding it } (128-bit CAS, ties to the earliest row).
// An aggregate over a GROUP BY column needs no state at all: inside one group that column has ONE value, so
// SUM = count * value and MIN = MAX = the value of the group's first row (agg_key[a] = the key part).
// expand_packed_kernel rewrites the occupied lines as general entries, so everything behind the scan is unchanged.
struct PackedLayout {
    int32_t entry_bytes;   // a multiple of 32 (whole sectors), at most 192; 0: this plan has no packed form
    int32_t id_words;      // words a lookup reads: 1 + key words
    int16_t key_word[4];   // first word of key part g
    int16_t key_wide[4];
    int16_t agg_off[4];    // byte offset of the state of lean aggregate a (index into l_agg[]); -1: derived from a key
    int16_t agg_key[4];    // key part the aggregate's column is (-1: none)
    int32_t count_off;     // byte offset of the row count (after expansion: the entry's index among the expanded ones)
    int32_t pad;
};

struct JoinSlot {
    uint64_t h;  // 0 empty, bit 63 lock
    uint64_t w0, w1;
    uint32_t tag;   // key tag; bit 31 (kJoinMatched): a left row matched this key (RIGHT / FULL joins)
    uint32_t head;  // index of the most recently added right row + 1 (0: none)
};
constexpr uint32_t kJoinMatched = 0x80000000u;
// okey of a right row without a match (RIGHT / FULL): behind every left row (their offsets stay below 2^45), in
// right-file order; the row has no left row, its right row is the offset itself
constexpr uint64_t kOkeyRightOnly = 1ull << 61;

struct DevPlan {
    // ---- left file ----
    const uint8_t* data;
    uint64_t size;        // bytes of the file view
    uint64_t own_lo;      // rows whose first byte is in [own_lo, own_hi) belong to this scan
    uint64_t own_hi;
    uint64_t global_base; // offset of data[0] in the whole (multi-GPU) file: added to row offsets in okeys
    uint8_t delim, quote;
    uint8_t exact_only;   // dialect the mask fast path does not cover (whitespace delimiter, exotic quote)
    uint8_t signed_hint;  // the head of the file holds a field that starts with a sign: COUNT-WHERE plans take the general scalar loop
    int32_t mode;         // ScanMode
    int32_t first_tile, n_tiles;

    // ---- column slots ----
    int32_t n_left_cols;              // query column >= n_left_cols addresses the right table
    int32_t n_cols_total;
    int32_t nwantL, nwantR;
    int16_t wantL[kMaxSlots];         // ascending CSV column indices
    int16_t wantR[kMaxSlots];
    int16_t colslot[kMaxQueryCols];   // query column -> slot (left slots first), -1 unused
    int16_t gap[kMaxSlots];           // delimiters between wanted left field k-1 and k (gap[0] = wantL[0])

    // ---- predicate ----
    int32_t pred_kind;  // 0 none, 1 fused (program below, in the parameter block), 2 generic interpreter
    int32_t n_fcode;
    FInsn fcode_inl[kMaxFused];
    int16_t frefs_inl[kMaxFusedRefs];
    DConst consts_inl[kMaxFusedConsts];
    DPred pred;  // generic program + string pool (global memory)

    // ---- lean kernel plans (cqg_lean.cuh): 1 = no GROUP BY, 2 = GROUP BY; 0 = general kernel only ----
    int32_t simple;
    int32_t lean_k;       // lean GROUP BY, few groups: try lean2k_kernel (cqg_lean2k.cuh) before lean2g_kernel
    // lean kernel WHERE: postfix program over up to kMaxLeanLeaf leaves. Leaf kinds:
    //   0  column <op> decimal literal   (mant * A[fd] vs LB[fd]; lop 0 >, 1 <, 2 ==, 3 !=)
    //   1  column =  'text'  /  2  column != 'text'   (text of 1..16 bytes, packed like a key part)
    // prog[i] >= 0: push leaf i; -1 AND, -2 OR, -3 NOT.
    int32_t l_nleaf, l_nprog;
    LeanLeaf l_leaf[kMaxLeanLeaf];
    int8_t l_prog[16];
    int32_t lean_global;  // lean GROUP BY updates the packed global table directly (any number of groups)
    int32_t l_nagg;       // aggregates with a state (SUM/AVG/MIN/MAX over a known column): at most 4 on the lean kernel
    int32_t l_agg[4];     // their indices in aggs[]
    int32_t crlf;         // the head of the file holds a CR: the lean kernels take '\r' as a line terminator (cqg_lean2g.cuh: l2g_masks16)
    const double* dec_table;  // [4][10000]: correctly rounded mant / 10^fd for mant < 10000 (lean MIN/MAX keys)
    // work the lean kernel hands to the general one
    int32_t* def_tiles;                 // tiles with bytes the lean kernel does not classify (CR, quotes, blanks, file edges)
    unsigned long long* def_tile_count;
    uint64_t* def_rows;                 // single rows (64 bytes or longer, fields that are not short decimals)
    unsigned long long* def_row_count;
    uint64_t def_row_cap;
    const int32_t* tile_list;           // general kernel: when set, tile i of this launch is tile_list[i]
    // packed table of the lean GROUP BY in global mode
    PackedLayout pk;
    uint8_t* ptab;
    uint64_t pcap;                      // power of two
    unsigned long long* pcount;         // occupied lines
    int32_t hc_debug;                   // CQG_HC_DEBUG (measurement only): 1 skip the updates, 2 skip the table altogether
    int32_t edge_in_kernel;             // the lean kernels process the tiles at the file's edges themselves (cqg_lean2.cuh: LeanEdge)

    // ---- aggregation ----
    int32_t ngc;
    int16_t gcol[CQG_MAX_GROUP_COLS];
    int16_t gslot[CQG_MAX_GROUP_COLS];  // field slots of the key columns (-1: NULL)
    int32_t naggs;
    AggSpec aggs[CQG_MAX_AGGS];
    int32_t entry_bytes;
    int32_t scalar_regs;   // no GROUP BY and <= 4 SUM/AVG aggregates: accumulate in registers
    uint8_t* gtab;         // global table: gcap entries
    uint64_t gcap;         // power of two
    unsigned long long* gcount;  // occupied entries
    const uint8_t* entry_init;   // image of an empty entry
    int32_t smem_cap;      // entries of the per-CTA table (power of two, 0: none)
    int32_t smem_table_off;  // byte offset of the table in dynamic shared memory

    // ---- join ----
    const uint8_t* rdata;  // right file bytes
    uint64_t rsize;
    int32_t join;          // 0 none, 1 probe
    int32_t join_type;     // cqg_join_type_t: LEFT / FULL emit left rows without a match, RIGHT / FULL right rows without one
    int32_t join_pad;
    int32_t jl_col, jr_col;   // key columns (query column index; jr_col relative to the right table)
    JoinSlot* jslots;
    uint64_t jcap;
    uint64_t* jrow_off;
    uint32_t* jrow_next;
    unsigned long long* jrow_count;
    uint64_t jrow_cap;
    unsigned* jclass;   // [2] OR of (1 << class) over left / right key values
    int32_t need_right_fields;

    // ---- partition (SCAN_PARTITION): this scan's row offsets split by owner = join-key hash % part_world ----
    int32_t part_world;
    int32_t part_pad;
    unsigned long long* part_counts;  // [world] rows per owner (first pass) / write cursors (second pass)
    const uint64_t* part_base;        // [world] first index of an owner's segment in part_list
    uint64_t* part_list;              // global row offsets; nullptr: count only

    // ---- select ----
    uint64_t* sel_okey;
    uint64_t* sel_roff;
    unsigned long long* sel_count;
    uint64_t sel_cap;

    // ---- bookkeeping ----
    unsigned* errflags;
    unsigned long long* rows_scanned;
};

enum ScanMode : int {
    SCAN_AGG = 0,
    SCAN_SELECT = 1,
    SCAN_COUNT_ROWS = 2,  // only rows_scanned
    SCAN_JOIN_BUILD = 3,  // insert (key, row offset) of every row into the join table
    SCAN_PARTITION = 4    // split the row offsets by owner of the join key (hash-partitioned join)
};

}  // namespace cqg
