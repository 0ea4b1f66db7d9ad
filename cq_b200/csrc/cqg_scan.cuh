// cqg_scan.cuh — the fused scan kernel of the cq hot path for sm_100a.
//
// One persistent CTA per slot walks tiles of the resident CSV bytes:
//   TMA bulk copy (cp.async.bulk + mbarrier) of tile + overlap into shared memory
//   -> phase 1 : SWAR byte classification, 16 bytes per thread: terminator / delimiter bitmasks
//   -> phase 1b: row starts = terminator->non-terminator transitions, block prefix sum, row list
//   -> phase 2 : one row per thread: field split on the bitmasks (exact sequential splitter
//                for rows with quote characters), typed decode, WHERE, then
//                aggregation (registers / shared-memory table / global table), join probe,
//                or selection.
// Reference loops replaced: csv_load line loop src/csv_reader.c:404-427, parse_line :278-338,
// parse_value :195-240, filter_rows evaluator_utils.c:986-1006, create_groups
// evaluator_aggregates.c:108-176, evaluate_aggregate :263-326, perform_join
// evaluator_joins.c:96-140.
#pragma once
#include "cqg_rtc.h"
#ifndef __CUDACC_RTC__
#include <cstdint>
#include <cuda_runtime.h>
#endif

#include "cqg_plan.cuh"

namespace cqg {

// ------------------------------------------------------------------------------------------
// geometry
// ------------------------------------------------------------------------------------------
template <int THREADS_, int TILE_, int STAGES_, int OVER_ = 992>
struct Geo {
    static constexpr int THREADS = THREADS_, TILE = TILE_, STAGES = STAGES_;
    static constexpr int PRE = 32;                 // bytes kept in front of the tile (>= 1 needed)
    static constexpr int OVER = OVER_;             // bytes after the tile a row may run into
    static constexpr int BUF = PRE + TILE + OVER;  // multiple of 32
    static constexpr int CHUNKS = BUF / 16;
    static constexpr int MASKW = BUF / 32 + 4;     // mask words incl. padding for 64-bit windows
    static constexpr int ROWCAP = 2048;            // rows listed per pass
    static constexpr int NWARPS = THREADS / 32;
    static constexpr int WPT = (TILE / 32) / THREADS;  // mask words per thread in phase 1b
    static_assert((TILE / 32) % THREADS == 0, "tile words must divide by threads");
    // shared memory layout (bytes)
    static constexpr int OFF_BUF = 0;
    static constexpr int OFF_TM = OFF_BUF + STAGES * BUF;
    static constexpr int OFF_DM = OFF_TM + MASKW * 4;
    static constexpr int OFF_QM = OFF_DM + MASKW * 4;
    static constexpr int OFF_ROW = OFF_QM + MASKW * 4;
    static constexpr int OFF_WSUM = OFF_ROW + ROWCAP * 2;
    static constexpr int OFF_MBAR = OFF_WSUM + 64 * 4;
    static constexpr int OFF_TABLE = (OFF_MBAR + STAGES * 8 + 127) / 128 * 128;
};

// ------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D TMA bulk copy (SASS: UBLKCP / SYNCS)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    uint32_t addr = smem_u32(bar);
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ------------------------------------------------------------------------------------------
// SWAR byte classification
// ------------------------------------------------------------------------------------------
// 0x80 in every byte of v that equals the byte replicated in pat
__device__ __forceinline__ uint32_t eq_flags(uint32_t v, uint32_t pat) {
    uint32_t t = v ^ pat;
    uint32_t a = (t & 0x7f7f7f7fu) + 0x7f7f7f7fu;
    return ~(a | t | 0x7f7f7f7fu);
}
// 0x80 flags -> 4-bit mask in byte order (no carries: the 16 partial-product bits are distinct)
__device__ __forceinline__ uint32_t flags_to_mask4(uint32_t f) { return (f * 0x00204081u) >> 28; }

__device__ __forceinline__ uint32_t eq_mask16(const uint4& v, uint32_t pat) {
    return flags_to_mask4(eq_flags(v.x, pat)) | (flags_to_mask4(eq_flags(v.y, pat)) << 4) |
           (flags_to_mask4(eq_flags(v.z, pat)) << 8) | (flags_to_mask4(eq_flags(v.w, pat)) << 12);
}

// 0x80 in every byte of v equal to the (ASCII) byte replicated in pat: 3 instructions
__device__ __forceinline__ uint32_t eq_flags7(uint32_t v, uint32_t pat) {
    uint32_t a = ((v ^ pat) & 0x7f7f7f7fu) + 0x7f7f7f7fu;
    return ~(a | v) & 0x80808080u;
}
// four flag words -> 16-bit mask in byte order, through the high half of four multiplies
__device__ __forceinline__ uint32_t flags_to_mask16(uint32_t f0, uint32_t f1, uint32_t f2, uint32_t f3) {
    uint32_t lo = (__umulhi(f0, 0x02040810u) & 0x0fu) | (__umulhi(f1, 0x20408100u) & 0xf0u);
    uint32_t hi = (__umulhi(f2, 0x02040810u) & 0x0fu) | (__umulhi(f3, 0x20408100u) & 0xf0u);
    return hi * 256u + lo;
}

// 64-bit window of a bitmask starting at bit `pos`
__device__ __forceinline__ unsigned long long mask_window(const uint32_t* m, uint32_t pos) {
    uint32_t w = pos >> 5, b = pos & 31u;
    uint32_t a0 = m[w], a1 = m[w + 1], a2 = m[w + 2];
    uint32_t lo = __funnelshift_r(a0, a1, b), hi = __funnelshift_r(a1, a2, b);
    return ((unsigned long long)hi << 32) | lo;
}

// ------------------------------------------------------------------------------------------
// ordered images for MIN/MAX
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t img_of_int(long long i) { return (uint64_t)i ^ 0x8000000000000000ull; }
__host__ __device__ __forceinline__ long long int_of_img(uint64_t u) { return (long long)(u ^ 0x8000000000000000ull); }
__host__ __device__ __forceinline__ uint64_t img_of_bits(uint64_t b) {
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ uint64_t bits_of_img(uint64_t u) {
    return (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
}

// ------------------------------------------------------------------------------------------
// canonical GROUP BY key parts (the reference's key string, evaluator_aggregates.c:121-141,
// as (tag, w0, w1) so that equal strings <=> equal triples)
// ------------------------------------------------------------------------------------------
// |x| * 10^6 rounded half-even on the exact binary value = the digits "%.6f" prints, as a
// 128-bit integer (w1:w0). Returns false for |x| >= 2^53 (an integer: the caller keys on the bits).
__device__ inline bool dbl_key6(double x, uint64_t& w0, uint64_t& w1) {
    uint64_t bits = (uint64_t)__double_as_longlong(x) & 0x7fffffffffffffffull;
    uint32_t e = (uint32_t)(bits >> 52);
    w0 = 0;
    w1 = 0;
    if (e == 0) return true;  // subnormal / zero: far below 5e-7
    uint64_t m = (bits & 0xfffffffffffffull) | (1ull << 52);
    int e2 = (int)e - 1075;  // |x| = m * 2^e2
    if (e2 >= 0) return false;
    uint64_t lo = m * 1000000ull, hi = __umul64hi(m, 1000000ull);  // P = m * 10^6 < 2^73
    int s = -e2;
    if (s >= 74) return true;  // P < half of 2^s
    uint64_t q_lo, q_hi, rem_hi, rem_lo, half_hi, half_lo;
    if (s < 64) {
        q_lo = (lo >> s) | (hi << (64 - s));
        q_hi = hi >> s;
        rem_hi = 0;
        rem_lo = lo & ((1ull << s) - 1ull);
        half_hi = 0;
        half_lo = 1ull << (s - 1);
    } else {
        int t = s - 64;
        q_lo = t ? (hi >> t) : hi;
        q_hi = 0;
        rem_hi = t ? (hi & ((1ull << t) - 1ull)) : 0;
        rem_lo = lo;
        half_hi = t ? (1ull << (t - 1)) : 0;
        half_lo = t ? 0 : (1ull << 63);
    }
    bool gt = rem_hi > half_hi || (rem_hi == half_hi && rem_lo > half_lo);
    bool eq = rem_hi == half_hi && rem_lo == half_lo;
    if (gt || (eq && (q_lo & 1ull))) {
        q_lo++;
        if (q_lo == 0ull) q_hi++;
    }
    w0 = q_lo;
    w1 = q_hi;
    return true;
}

__device__ __forceinline__ void hash_str2(const uint8_t* p, uint32_t n, uint64_t& h0, uint64_t& h1) {
    uint64_t a = 0x9E3779B97F4A7C15ull ^ n, b = 0xC2B2AE3D27D4EB4Full + n;
    for (uint32_t k = 0; k < n; k++) {
        uint64_t c = p[k];
        a = (a ^ c) * 0x100000001B3ull;
        b = (b + c) * 0x9FB21C651E98DF25ull;
        b ^= b >> 29;
    }
    h0 = mix64(a);
    h1 = mix64(b ^ 0x5555555555555555ull);
}

__device__ __forceinline__ void pack_str16(const uint8_t* p, uint32_t n, uint64_t& w0, uint64_t& w1) {
    w0 = 0;
    w1 = 0;
    for (uint32_t k = 0; k < n && k < 8; k++) w0 |= (uint64_t)p[k] << (8 * k);
    for (uint32_t k = 8; k < n; k++) w1 |= (uint64_t)p[k] << (8 * (k - 8));
}

// group_mode: reference key-string identity (NULL == "NULL", strings cut at 255 bytes, %.6f doubles).
// join mode : value_compare identity inside one comparison class (numeric as double).
template <bool GROUP>
__device__ inline void canon_part(const DVal& v, bool multi, unsigned& err, uint32_t& tag, uint64_t& w0, uint64_t& w1) {
    w0 = 0;
    w1 = 0;
    switch (v.type) {
        case T_INT:
            if (GROUP) {
                tag = KT_INT;
                w0 = (uint64_t)v.i;
            } else {
                double d = (double)v.i;  // value_compare compares numerics as doubles (csv_reader.c:116-121)
                tag = KT_DBL_POS;
                w0 = (uint64_t)__double_as_longlong(d + 0.0);
            }
            break;
        case T_DBL:
            if (GROUP) {
                tag = (__double_as_longlong(v.d) < 0) ? KT_DBL_NEG : KT_DBL_POS;
                if (!dbl_key6(v.d, w0, w1)) {  // |x| >= 2^53: "%.6f" prints the integer itself
                    tag = KT_DBL_BIG;
                    w0 = (uint64_t)__double_as_longlong(v.d);
                }
            } else {
                tag = KT_DBL_POS;
                double d = v.d == 0.0 ? 0.0 : v.d;  // -0.0 == 0.0
                w0 = (uint64_t)__double_as_longlong(d);
            }
            break;
        case T_DATE:
            tag = KT_DATE;
            w0 = (uint64_t)v.i;
            break;
        case T_STR: {
            uint32_t n = v.len;
            if (GROUP && n > 255u) n = 255u;  // strncpy(key, s, 255)
            if (GROUP && n == 4 && v.s[0] == 'N' && v.s[1] == 'U' && v.s[2] == 'L' && v.s[3] == 'L') {
                tag = KT_NULL;
                break;
            }
            if (GROUP && n == 10 && v.s[4] == '-' && v.s[7] == '-') {
                // a text that reads exactly like a rendered date ("%04d-%02d-%02d") shares the DATE's key string;
                // it is text only because blanks made the raw field longer than the 10 bytes parse_date is tried on
                bool dig = true;
                for (int k = 0; k < 10; k++)
                    if (k != 4 && k != 7) dig = dig && is_digit(v.s[k]);
                if (dig) {
                    long long y = (v.s[0] - 48) * 1000 + (v.s[1] - 48) * 100 + (v.s[2] - 48) * 10 + (v.s[3] - 48);
                    long long mo = (v.s[5] - 48) * 10 + (v.s[6] - 48), d = (v.s[8] - 48) * 10 + (v.s[9] - 48);
                    if (valid_date(y, mo, d)) {
                        tag = KT_DATE;
                        w0 = (uint64_t)((y << 16) | (mo << 8) | d);
                        break;
                    }
                }
            }
            if (GROUP && multi) {
                for (uint32_t k = 0; k < n; k++)
                    if (v.s[k] == '\t') err |= KERR_KEY_TAB;
            }
            if (n <= 16u) {
                tag = KT_STR;
                pack_str16(v.s, n, w0, w1);
            } else {
                tag = KT_STR_HASH;
                hash_str2(v.s, n, w0, w1);
            }
            break;
        }
        default:
            tag = KT_NULL;
            break;
    }
}

__device__ __forceinline__ uint64_t key_hash_step(uint64_t h, uint32_t tag, uint64_t w0, uint64_t w1) {
    h = (h ^ (w0 + tag)) * 0xFF51AFD7ED558CCDull;
    h ^= h >> 32;
    h = (h ^ w1) * 0xC4CEB9FE1A85EC53ull;
    h ^= h >> 29;
    return h;
}
__device__ __forceinline__ uint64_t key_hash_final(uint64_t h) {
    h = mix64(h);
    return (h & 0x7fffffffffffffffull) | 1ull;
}

// ------------------------------------------------------------------------------------------
// group table: find-or-insert (open addressing, linear probing, 64-bit hash tag per slot)
// ------------------------------------------------------------------------------------------
template <bool SM>
__device__ __forceinline__ uint64_t load_u64(const void* p) {
    return *(const volatile uint64_t*)p;
}

__device__ __forceinline__ void copy_entry_init(uint8_t* e, const uint8_t* init, int bytes) {
    // entry_bytes is a multiple of 8; the hash word (offset 0) is written by the caller
    for (int o = 8; o < bytes; o += 8) *(uint64_t*)(e + o) = *(const uint64_t*)(init + o);
}

// returns the entry, or nullptr when the table is full. `fresh_init`: global entries are
// pre-initialised by the host; shared entries by the CTA prologue — so insertion only writes keys.
__device__ inline uint8_t* table_find_insert(uint8_t* tab, uint64_t cap, int entry_bytes, int ngc, uint64_t h, uint32_t tags,
                                             const uint64_t* kw /* [ngc][2] */, unsigned long long* occupancy,
                                             uint64_t max_occupancy) {
    uint64_t mask = cap - 1;
    uint64_t i = (h >> 1) & mask;
    for (uint64_t probes = 0; probes < cap;) {
        uint8_t* e = tab + i * (uint64_t)entry_bytes;
        unsigned long long* hp = (unsigned long long*)(e + kOffHash);
        unsigned long long cur = *(volatile unsigned long long*)hp;
        if (cur == 0ull) {
            if (occupancy && *(volatile unsigned long long*)occupancy >= max_occupancy) return nullptr;
            cur = atomicCAS(hp, 0ull, (unsigned long long)(h | kLockBit));
            if (cur == 0ull) {
                if (occupancy) atomicAdd(occupancy, 1ull);
                *(uint32_t*)(e + kOffTags) = tags;
                for (int g = 0; g < ngc; g++) {
                    *(uint64_t*)(e + kOffKeys + 16 * g) = kw[2 * g];
                    *(uint64_t*)(e + kOffKeys + 16 * g + 8) = kw[2 * g + 1];
                }
                __threadfence();
                atomicExch(hp, (unsigned long long)h);
                return e;
            }
        }
        if ((cur & ~kLockBit) == h) {
            if (cur & kLockBit) continue;  // being initialised by another thread: look again
            __threadfence();
            bool same = *(volatile uint32_t*)(e + kOffTags) == tags;
            for (int g = 0; g < ngc && same; g++)
                same = *(volatile uint64_t*)(e + kOffKeys + 16 * g) == kw[2 * g] &&
                       *(volatile uint64_t*)(e + kOffKeys + 16 * g + 8) == kw[2 * g + 1];
            if (same) return e;
        }
        i = (i + 1) & mask;
        probes++;
    }
    return nullptr;
}

// ------------------------------------------------------------------------------------------
// aggregate state updates
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void amin64(uint64_t* p, uint64_t v) {
    if (v < *(volatile uint64_t*)p) atomicMin((unsigned long long*)p, (unsigned long long)v);
}
__device__ __forceinline__ void amax64(uint64_t* p, uint64_t v) {
    if (v > *(volatile uint64_t*)p) atomicMax((unsigned long long*)p, (unsigned long long)v);
}

__device__ __forceinline__ const uint8_t* str_ref_ptr(const DevPlan& P, uint64_t ref, uint32_t& len) {
    len = (uint32_t)(ref & 0x3ffffu);
    uint64_t off = (ref >> 18) & 0x1fffffffffffull;
    return ((ref >> 63) ? P.rdata : P.data) + off;
}

// keep the smaller (MIN) / larger (MAX) string; ties keep the incumbent
__device__ inline void str_extreme(const DevPlan& P, uint64_t* slot, uint64_t cand, bool is_min) {
    const uint64_t empty = is_min ? ~0ull : 0ull;
    uint32_t cl;
    const uint8_t* cp = str_ref_ptr(P, cand, cl);
    for (;;) {
        uint64_t cur = *(volatile uint64_t*)slot;
        if (cur != empty) {
            uint32_t ul;
            const uint8_t* up = str_ref_ptr(P, cur, ul);
            int c = str_cmp(cp, cl, up, ul);
            if (is_min ? c >= 0 : c <= 0) return;
        }
        if (atomicCAS((unsigned long long*)slot, (unsigned long long)cur, (unsigned long long)cand) == cur) return;
    }
}

// 128-bit compare-and-swap (atom.cas.b128, sm_90+): global or shared address, 16-byte aligned
__device__ __forceinline__ bool cas128(void* addr, uint64_t& exp_lo, uint64_t& exp_hi, uint64_t new_lo, uint64_t new_hi) {
    uint64_t old_lo, old_hi;
    if (__isShared(addr)) {
        asm volatile(
            "{\n\t.reg .b128 c, s, d;\n\t"
            "mov.b128 c, {%2, %3};\n\t"
            "mov.b128 s, {%4, %5};\n\t"
            "atom.shared.cas.b128 d, [%6], c, s;\n\t"
            "mov.b128 {%0, %1}, d;\n\t}"
            : "=l"(old_lo), "=l"(old_hi)
            : "l"(exp_lo), "l"(exp_hi), "l"(new_lo), "l"(new_hi), "r"(smem_u32(addr))
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .b128 c, s, d;\n\t"
            "mov.b128 c, {%2, %3};\n\t"
            "mov.b128 s, {%4, %5};\n\t"
            "atom.global.cas.b128 d, [%6], c, s;\n\t"
            "mov.b128 {%0, %1}, d;\n\t}"
            : "=l"(old_lo), "=l"(old_hi)
            : "l"(exp_lo), "l"(exp_hi), "l"(new_lo), "l"(new_hi), "l"(addr)
            : "memory");
    }
    bool ok = old_lo == exp_lo && old_hi == exp_hi;
    exp_lo = old_lo;
    exp_hi = old_hi;
    return ok;
}

// numeric extreme: (key, okey) where key orders the values as value_compare does (as doubles,
// -0.0 == 0.0) and okey breaks ties towards the earliest row — "keep the first value unless a
// later one is strictly less/greater" (evaluator_aggregates.c:316-321). The value itself (its
// type and bits) is re-read from that row when the result is built.
__device__ __forceinline__ void num_extreme(uint64_t* st /*16B aligned*/, uint64_t key, uint64_t okey, bool is_min) {
    uint64_t ck = *(volatile uint64_t*)&st[0];
    uint64_t co = *(volatile uint64_t*)&st[1];
    for (;;) {
        bool better = is_min ? (key < ck || (key == ck && okey < co)) : (key > ck || (key == ck && okey < co));
        if (!better) return;
        if (cas128(st, ck, co, key, okey)) return;
    }
}

__device__ __forceinline__ uint64_t num_key(double d) {
    d = d + 0.0;  // -0.0 -> +0.0
    return img_of_bits((uint64_t)__double_as_longlong(d));
}

// MIN/MAX state: [0] first_nonnull (okey<<2|class)  [1] date  [2,3] numeric (key, okey)  [4] string ref  [5] pad
// one value enters it (evaluator_aggregates.c:311-326, order-free form)
__device__ inline void minmax_update(const DevPlan& P, uint8_t* st, bool is_min, const DVal& v, uint64_t okey, bool right_side,
                                     const uint8_t* field_base_file, unsigned& err) {
    if (v.type == T_NULL) return;
    uint64_t* s = (uint64_t*)st;
    uint32_t cls = (v.type == T_INT || v.type == T_DBL) ? 1u : (v.type == T_STR ? 2u : 3u);
    amin64(&s[0], (okey << 2) | cls);
    if (v.type == T_INT || v.type == T_DBL) {
        num_extreme(&s[2], num_key(v.type == T_INT ? (double)v.i : v.d), okey, is_min);
    } else if (v.type == T_DATE) {
        uint64_t im = (uint64_t)v.i + 1ull;  // keep 0 / ~0 free as "empty"
        if (is_min) amin64(&s[1], im); else amax64(&s[1], im);
    } else {
        // string bytes live in the resident file: reference them
        uint64_t off = (uint64_t)(v.s - field_base_file);
        if (v.len > 0x3ffffu) {
            err |= KERR_STR_LONG;
            return;
        }
        if (off >> 45) {
            err |= KERR_OFFSET_RANGE;
            return;
        }
        uint64_t ref = (right_side ? (1ull << 63) : 0ull) | (off << 18) | v.len;
        str_extreme(P, &s[4], ref, is_min);
    }
}

// merge MIN/MAX state `src` into `dst` (shared -> global flush, partial merge)
__device__ inline void minmax_merge(const DevPlan& P, uint8_t* dst, const uint8_t* src, bool is_min) {
    uint64_t* d = (uint64_t*)dst;
    const uint64_t* s = (const uint64_t*)src;
    amin64(&d[0], s[0]);
    const uint64_t empty = is_min ? ~0ull : 0ull;
    if (s[1] != empty) {
        if (is_min) amin64(&d[1], s[1]); else amax64(&d[1], s[1]);
    }
    if (s[2] != empty) num_extreme(&d[2], s[2], s[3], is_min);
    if (s[4] != empty) str_extreme(P, &d[4], s[4], is_min);
}

// ------------------------------------------------------------------------------------------
// per-thread register accumulators for the no-GROUP-BY case
// ------------------------------------------------------------------------------------------
struct ThreadAcc {
    uint32_t rows;     // data rows seen
    uint32_t count;    // rows passing WHERE
    uint64_t first;    // min okey
    long long si[4];   // INTEGER-typed values
    double sd[4];      // DOUBLE-typed values
    long long s3[4];   // short decimals, summed exactly as value * 1000 (simple route)
    uint32_t sn[4];
    unsigned err;
};

// ------------------------------------------------------------------------------------------
// rows as the operators see them
// ------------------------------------------------------------------------------------------
// FastRow<NW>: a row of a tile staged in shared memory whose wanted fields sit in registers
// (NW <= 4 keeps off[]/len[] out of local memory). No join. `clean` = the tile has no byte
// below 0x23 except '\n', which licenses decode_field_clean.
template <int NW>
struct FastRow {
    const uint8_t* base;   // shared-memory tile
    const uint8_t* lfile;  // base - (file offset of base[0])
    uint32_t off[NW], len[NW];
    bool clean;
    static constexpr bool kJoined = false;

    __device__ __forceinline__ DVal value(const DevPlan& P, int col, unsigned& err) const {
        DVal v;
        v.type = T_NULL;
        v.len = 0;
        v.i = 0;
        if (col < 0 || col >= P.n_cols_total) return v;
        const int s = P.colslot[col];
        if (s < 0) return v;
        uint32_t o, l;
        if (NW <= 4) {
            o = off[0];
            l = len[0];
#pragma unroll
            for (int k = 1; k < NW; k++)
                if (s == k) {
                    o = off[k];
                    l = len[k];
                }
        } else {
            o = off[s];
            l = len[s];
        }
        return clean ? decode_field_clean(base + o, l, err) : decode_field(base + o, l, err);
    }
    // field by slot (the host resolves columns to slots once per plan)
    __device__ __forceinline__ DVal value_slot(const DevPlan&, int s, unsigned& err) const {
        if (s < 0) {
            DVal v;
            v.type = T_NULL;
            v.len = 0;
            v.i = 0;
            return v;
        }
        uint32_t o, l;
        if (NW <= 4) {
            o = off[0];
            l = len[0];
#pragma unroll
            for (int k = 1; k < NW; k++)
                if (s == k) {
                    o = off[k];
                    l = len[k];
                }
        } else {
            o = off[s];
            l = len[s];
        }
        return clean ? decode_field_clean(base + o, l, err) : decode_field(base + o, l, err);
    }
    __device__ __forceinline__ const uint8_t* file_base(bool) const { return lfile; }
};

// SlowRow: the general case (joined rows, rows read straight from HBM, generic predicate)
struct SlowRow {
    RowView rv;
    static constexpr bool kJoined = true;
    __device__ __forceinline__ DVal value(const DevPlan&, int col, unsigned& err) const { return row_value(rv, col, err); }
    __device__ __forceinline__ DVal value_slot(const DevPlan&, int s, unsigned& err) const {
        if (s < 0) {
            DVal v;
            v.type = T_NULL;
            v.len = 0;
            v.i = 0;
            return v;
        }
        const uint8_t* b = s < rv.nleft_slots ? rv.base : rv.rbase;
        return decode_field(b + rv.foff[s], rv.flen[s], err);
    }
    __device__ __forceinline__ const uint8_t* file_base(bool right) const { return right ? rv.rfile : rv.lfile; }
};

// ------------------------------------------------------------------------------------------
// predicate evaluation
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ DVal const_value(const DevPlan& P, int idx) {
    const DConst c = P.consts_inl[idx];
    DVal v;
    v.type = c.type;
    v.len = c.len;
    v.i = c.bits;
    if (c.type == T_STR) v.s = P.pred.pool + c.bits;
    return v;
}

template <class Row>
__device__ __forceinline__ DVal fetch_ref(const DevPlan& P, const Row& row, int ref, unsigned& err) {
    if (ref == kRefNull) {
        DVal v;
        v.type = T_NULL;
        v.len = 0;
        v.i = 0;
        return v;
    }
    if (ref & kRefConst) return const_value(P, ref & 0x1fff);
    return row.value_slot(P, ref & 0x1fff, err);
}

// the program lives in the kernel parameter block (constant bank): no dependent global loads
template <class Row>
__device__ __forceinline__ bool leaf_cmp(const DevPlan& P, const Row& row, const FInsn in, unsigned& err) {
    DVal l = fetch_ref(P, row, in.a, err), r = fetch_ref(P, row, in.b, err);
    int c;
    if (l.type == T_INT && r.type == T_INT && ((unsigned long long)(l.i + (1ll << 52)) >> 53) == 0 &&
        ((unsigned long long)(r.i + (1ll << 52)) >> 53) == 0)
        c = l.i < r.i ? -1 : (l.i > r.i ? 1 : 0);  // same answer as the double compare below 2^52
    else
        c = val_compare(l, r);
    return in.n == CQG_OP_EQ ? c == 0 : in.n == CQG_OP_NE ? c != 0 : in.n == CQG_OP_GT ? c > 0
         : in.n == CQG_OP_LT ? c < 0 : in.n == CQG_OP_GE ? c >= 0 : c <= 0;
}

template <class Row>
__device__ __forceinline__ bool eval_fused(const DevPlan& P, const Row& row, unsigned& err) {
    if (P.n_fcode == 1 && P.fcode_inl[0].op == F_CMP) return leaf_cmp(P, row, P.fcode_inl[0], err);  // `col op literal`
    uint32_t bs = 0;  // bit stack, top = bit 0
    for (int pc = 0; pc < P.n_fcode; pc++) {
        const FInsn in = P.fcode_inl[pc];
        bool b;
        switch (in.op) {
            case F_CMP: {
                b = leaf_cmp(P, row, in, err);
                bs = (bs << 1) | (uint32_t)b;
                break;
            }
            case F_IN: case F_NOT_IN: {
                DVal l = fetch_ref(P, row, in.a, err);
                bool found = false;
                for (int k = 0; k < in.n && !found; k++) {
                    DVal it = fetch_ref(P, row, P.frefs_inl[in.b + k], err);
                    found = val_compare(l, it) == 0;
                }
                b = in.op == F_IN ? found : !found;
                bs = (bs << 1) | (uint32_t)b;
                break;
            }
            case F_LIKE: case F_ILIKE: {
                DVal l = fetch_ref(P, row, in.a, err), r = fetch_ref(P, row, in.b, err);
                b = (l.type == T_STR && r.type == T_STR) ? like_match(l.s, l.len, r.s, r.len, in.op == F_LIKE) : false;
                bs = (bs << 1) | (uint32_t)b;
                break;
            }
            case F_AND: bs = ((bs >> 1) & ~1u) | ((bs >> 1) & bs & 1u); break;
            case F_OR: bs = ((bs >> 1) & ~1u) | (((bs >> 1) | bs) & 1u); break;
            case F_NOT: bs ^= 1u; break;
            case F_TRUE: bs = (bs << 1) | 1u; break;
            default: bs = bs << 1; break;  // F_FALSE
        }
    }
    return (bs & 1u) != 0;
}

// ------------------------------------------------------------------------------------------
// what happens to one row that passed WHERE
// ------------------------------------------------------------------------------------------
struct CtaState {
    uint8_t* stab;  // shared-memory group table (smem_cap entries) or nullptr
    unsigned long long* s_occ;  // its occupancy counter (shared)
};

// 64-bit integer add in shared memory with native 32-bit atomics: the lane that wraps the low word adds the carry
__device__ __forceinline__ void smem_add64(void* p, unsigned long long v) {
    const uint32_t vlo = (uint32_t)v, vhi = (uint32_t)(v >> 32);
    const uint32_t old = atomicAdd((unsigned int*)p, vlo);
    const uint32_t up = vhi + ((old + vlo) < old ? 1u : 0u);
    if (up) atomicAdd((unsigned int*)p + 1, up);
}

template <bool SM, class Row>
__device__ __forceinline__ void entry_accumulate(const DevPlan& P, uint8_t* e, const Row& row, uint64_t okey, unsigned& err) {
    amin64((uint64_t*)(e + kOffFirst), okey);
    if (SM) atomicAdd((unsigned int*)(e + kOffCount), 1u);
    else atomicAdd((unsigned long long*)(e + kOffCount), 1ull);
    for (int a = 0; a < P.naggs; a++) {
        const AggSpec sp = P.aggs[a];
        if (sp.off < 0) continue;
        DVal v = row.value_slot(P, sp.slot, err);
        uint8_t* st = e + sp.off;
        if (sp.func == CQG_AGG_SUM || sp.func == CQG_AGG_AVG) {
            if (v.type == T_INT && (unsigned long long)(v.i + (1ll << 31)) >> 32 == 0) {
                if (SM) smem_add64(st, (unsigned long long)v.i);
                else atomicAdd((unsigned long long*)st, (unsigned long long)v.i);  // exact, cannot overflow below 2^31 rows
                if (SM) atomicAdd((unsigned int*)(st + 16), 1u);
                else atomicAdd((unsigned long long*)(st + 16), 1ull);
            } else if (v.type == T_INT) {
                atomicAdd((double*)(st + 8), (double)v.i);
                if (SM) atomicAdd((unsigned int*)(st + 16), 1u);
                else atomicAdd((unsigned long long*)(st + 16), 1ull);
            } else if (v.type == T_DBL) {
                atomicAdd((double*)(st + 8), v.d);
                if (SM) atomicAdd((unsigned int*)(st + 16), 1u);
                else atomicAdd((unsigned long long*)(st + 16), 1ull);
            }
        } else {
            bool right = sp.col >= P.n_left_cols;
            minmax_update(P, st, sp.func == CQG_AGG_MIN, v, okey, right, row.file_base(right), err);
        }
    }
}

// fold entry `src` (same layout) into `dst`
__device__ __noinline__ void entry_merge(const DevPlan& P, uint8_t* dst, const uint8_t* src) {
    amin64((uint64_t*)(dst + kOffFirst), *(const uint64_t*)(src + kOffFirst));
    atomicAdd((unsigned long long*)(dst + kOffCount), (unsigned long long)*(const uint64_t*)(src + kOffCount));
    for (int a = 0; a < P.naggs; a++) {
        const AggSpec sp = P.aggs[a];
        if (sp.off < 0) continue;
        if (sp.func == CQG_AGG_SUM || sp.func == CQG_AGG_AVG) {
            unsigned long long si = *(const unsigned long long*)(src + sp.off);
            double sd = *(const double*)(src + sp.off + 8);
            unsigned long long n = *(const unsigned long long*)(src + sp.off + 16);
            unsigned long long s3 = *(const unsigned long long*)(src + sp.off + 24);
            if (si) atomicAdd((unsigned long long*)(dst + sp.off), si);
            if (sd != 0.0) atomicAdd((double*)(dst + sp.off + 8), sd);
            if (n) atomicAdd((unsigned long long*)(dst + sp.off + 16), n);
            if (s3) atomicAdd((unsigned long long*)(dst + sp.off + 24), s3);
        } else {
            minmax_merge(P, dst + sp.off, src + sp.off, sp.func == CQG_AGG_MIN);
        }
    }
}

__device__ __noinline__ uint8_t* global_entry_for(const DevPlan& P, uint64_t h, uint32_t tags, const uint64_t* kw, unsigned& err) {
    uint8_t* e = table_find_insert(P.gtab, P.gcap, P.entry_bytes, P.ngc, h, tags, kw, P.gcount, P.gcap / 2);
    if (!e) err |= KERR_TABLE_FULL;
    return e;
}

template <class Row>
__device__ __forceinline__ void agg_row(const DevPlan& P, const CtaState& cs, const Row& row, uint64_t okey, ThreadAcc& acc) {
    if (P.scalar_regs) {
        acc.count++;
        if (okey < acc.first) acc.first = okey;
#pragma unroll
        for (int a = 0; a < 4; a++) {
            if (a < P.naggs) {
                const AggSpec sp = P.aggs[a];
                if (sp.off >= 0) {
                    DVal v = row.value_slot(P, sp.slot, acc.err);
                    if (sp.func == CQG_AGG_SUM || sp.func == CQG_AGG_AVG) {
                        if (v.type == T_INT && (unsigned long long)(v.i + (1ll << 31)) >> 32 == 0) {
                            acc.si[a] += v.i;
                            acc.sn[a]++;
                        } else if (v.type == T_INT) {
                            acc.sd[a] += (double)v.i;
                            acc.sn[a]++;
                        } else if (v.type == T_DBL) {
                            acc.sd[a] += v.d;
                            acc.sn[a]++;
                        }
                    } else {
                        bool right = sp.col >= P.n_left_cols;
                        minmax_update(P, cs.stab + sp.off, sp.func == CQG_AGG_MIN, v, okey, right, row.file_base(right), acc.err);
                    }
                }
            }
        }
        return;
    }
    uint64_t kw[2 * CQG_MAX_GROUP_COLS];
    uint32_t tags = 0;
    uint64_t h = 0x243F6A8885A308D3ull + (uint64_t)P.ngc;
    for (int g = 0; g < P.ngc; g++) {
        DVal v = row.value_slot(P, P.gslot[g], acc.err);
        uint32_t tag;
        canon_part<true>(v, P.ngc > 1, acc.err, tag, kw[2 * g], kw[2 * g + 1]);
        tags |= tag << (4 * g);
        h = key_hash_step(h, tag, kw[2 * g], kw[2 * g + 1]);
    }
    h = key_hash_final(h);
    if (cs.stab) {
        uint8_t* e = table_find_insert(cs.stab, (uint64_t)P.smem_cap, P.entry_bytes, P.ngc, h, tags, kw, cs.s_occ,
                                       (uint64_t)(P.smem_cap - (P.smem_cap >> 2)));
        if (e) {
            entry_accumulate<true>(P, e, row, okey, acc.err);
            return;
        }
    }
    uint8_t* e = global_entry_for(P, h, tags, kw, acc.err);
    if (e) entry_accumulate<false>(P, e, row, okey, acc.err);
}

__device__ __forceinline__ void select_row(const DevPlan& P, uint64_t okey, uint64_t roff) {
    unsigned long long idx = atomicAdd(P.sel_count, 1ull);
    if (idx < P.sel_cap) {
        P.sel_okey[idx] = okey;
        if (P.sel_roff) P.sel_roff[idx] = roff;
    }
}

// ------------------------------------------------------------------------------------------
// join: build and probe (perform_join, evaluator_joins.c:96-140, as a hash equi-join)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t join_class(const DVal& v) {
    return v.type == T_NULL ? 0u : ((v.type == T_INT || v.type == T_DBL) ? 1u : (v.type == T_STR ? 2u : 3u));
}

__device__ inline JoinSlot* join_find(const DevPlan& P, uint64_t h, uint32_t tag, uint64_t w0, uint64_t w1, bool insert) {
    uint64_t mask = P.jcap - 1;
    uint64_t i = (h >> 1) & mask;
    for (uint64_t probes = 0; probes < P.jcap;) {
        JoinSlot* s = &P.jslots[i];
        unsigned long long cur = *(volatile unsigned long long*)&s->h;
        if (cur == 0ull) {
            if (!insert) return nullptr;
            cur = atomicCAS((unsigned long long*)&s->h, 0ull, (unsigned long long)(h | kLockBit));
            if (cur == 0ull) {
                s->w0 = w0;
                s->w1 = w1;
                s->tag = tag;
                __threadfence();
                atomicExch((unsigned long long*)&s->h, (unsigned long long)h);
                return s;
            }
        }
        if ((cur & ~kLockBit) == h) {
            if (cur & kLockBit) continue;
            __threadfence();
            if ((*(volatile uint32_t*)&s->tag & ~kJoinMatched) == tag && *(volatile uint64_t*)&s->w0 == w0 &&
                *(volatile uint64_t*)&s->w1 == w1)
                return s;
        }
        i = (i + 1) & mask;
        probes++;
    }
    return nullptr;
}

__device__ inline void join_key_of(const DVal& v, unsigned& err, uint32_t& tag, uint64_t& w0, uint64_t& w1, uint64_t& h) {
    canon_part<false>(v, false, err, tag, w0, w1);
    h = key_hash_final(key_hash_step(0x13198A2E03707344ull, tag, w0, w1));
}

__device__ inline void join_build_row(const DevPlan& P, const RowView& rv, uint64_t roff, unsigned& err) {
    DVal v = row_value(rv, P.jr_col, err);
    atomicOr(&P.jclass[1], 1u << join_class(v));
    uint32_t tag;
    uint64_t w0, w1, h;
    join_key_of(v, err, tag, w0, w1, h);
    JoinSlot* s = join_find(P, h, tag, w0, w1, true);
    if (!s) {
        err |= KERR_TABLE_FULL;
        return;
    }
    unsigned long long idx = atomicAdd(P.jrow_count, 1ull);
    if (idx >= P.jrow_cap) {
        err |= KERR_TABLE_FULL;
        return;
    }
    P.jrow_off[idx] = roff;
    P.jrow_next[idx] = atomicExch(&s->head, (uint32_t)idx + 1u);
}

// owner rank of a join key in a hash-partitioned join: bits of the hash the slot index does not use first
__device__ __forceinline__ uint32_t join_owner(uint64_t h, int world) { return (uint32_t)((h >> 33) % (uint64_t)world); }

// SCAN_PARTITION: the row's offset goes on the list of the rank that owns its key (both sides of the join
// canonicalise and hash the key the same way, so equal keys meet on one rank; NULL = NULL included)
__device__ inline void partition_row(const DevPlan& P, const RowView& rv, uint64_t goff, unsigned& err) {
    DVal v = row_value(rv, P.jr_col, err);
    uint32_t tag;
    uint64_t w0, w1, h;
    join_key_of(v, err, tag, w0, w1, h);
    // comparison classes of the keys this shard holds (first pass only): the caller ORs them over all ranks and both
    // sides, since a stray key of another class may land on a rank that owns no other key of the join
    if (!P.part_list) {
        const unsigned bit = 1u << join_class(v);
        if (!(*(volatile unsigned*)&P.jclass[0] & bit)) atomicOr(&P.jclass[0], bit);
    }
    const uint32_t o = join_owner(h, P.part_world);
    const unsigned long long idx = atomicAdd(&P.part_counts[o], 1ull);
    if (P.part_list) P.part_list[P.part_base[o] + idx] = P.global_base + goff;
}

// right row at file offset roff: find its end, split the wanted right columns
__device__ inline void split_right_row(const DevPlan& P, uint64_t roff, uint32_t* foff, uint32_t* flen) {
    const uint8_t* b = P.rdata + roff;
    uint64_t maxlen = P.rsize - roff;
    uint32_t re = 0;
    while ((uint64_t)re < maxlen && b[re] != '\n' && b[re] != '\r' && re < 0x7fffffffu) re++;
    split_row_exact(b, 0, re, P.delim, P.quote, P.wantR, P.nwantR, foff + P.nwantL, flen + P.nwantL);
}

// ------------------------------------------------------------------------------------------
// one data row, general route: `base` + foff/flen hold the wanted left fields; goff = file
// offset of the row. Kept out of line: the hot loop only pays for it when it is taken.
// ------------------------------------------------------------------------------------------
__device__ __noinline__ void process_row_slow(const DevPlan& P, const CtaState& cs, const uint8_t* base, const uint8_t* lfile,
                                              uint32_t* foff, uint32_t* flen, uint64_t goff, ThreadAcc& acc) {
    SlowRow row;
    RowView& rv = row.rv;
    rv.base = base;
    rv.rbase = nullptr;
    rv.lfile = lfile;
    rv.rfile = P.rdata;
    rv.foff = foff;
    rv.flen = flen;
    rv.colslot = P.colslot;
    rv.ncols_total = P.n_cols_total;
    rv.nleft_slots = P.nwantL;
    uint64_t gabs = P.global_base + goff;
    if (gabs >> 45) acc.err |= KERR_OFFSET_RANGE;
    uint64_t okey = gabs << 16;
    if (P.mode == SCAN_JOIN_BUILD) {
        join_build_row(P, rv, goff, acc.err);
        return;
    }
    if (P.mode == SCAN_PARTITION) {
        partition_row(P, rv, goff, acc.err);
        return;
    }
    if (!P.join) {
        bool pass = P.pred_kind == 0 ? true : (P.pred_kind == 1 ? eval_fused(P, row, acc.err) : eval_pred(P.pred, rv, acc.err));
        if (!pass) return;
        if (P.mode == SCAN_AGG) agg_row(P, cs, row, okey, acc);
        else select_row(P, okey, 0);
        return;
    }
    // probe (perform_join, evaluator_joins.c:96-139): every right row with an equal key, in right-file order; a left
    // row without a match comes out once with NULL right columns in LEFT / FULL joins
    const bool keep_unmatched_left = P.join_type == CQG_JOIN_LEFT || P.join_type == CQG_JOIN_FULL;
    JoinSlot* s = nullptr;
    if (P.jl_col >= 0 && P.jr_col >= 0) {  // else resolve_column -> NULL: the condition is false for every pair (joins.c:54)
        DVal lk = row_value(rv, P.jl_col, acc.err);
        atomicOr(&P.jclass[0], 1u << join_class(lk));
        uint32_t tag;
        uint64_t w0, w1, h;
        join_key_of(lk, acc.err, tag, w0, w1, h);
        s = join_find(P, h, tag, w0, w1, false);
    }
    if (!s || s->head == 0u) {
        if (!keep_unmatched_left) return;
        for (int k = 0; k < P.nwantR; k++) {  // every right column reads as NULL
            foff[P.nwantL + k] = 0;
            flen[P.nwantL + k] = 0;
        }
        rv.rbase = P.rdata;
        bool pass = P.pred_kind == 0 ? true : (P.pred_kind == 1 ? eval_fused(P, row, acc.err) : eval_pred(P.pred, rv, acc.err));
        if (!pass) return;
        if (P.mode == SCAN_AGG) agg_row(P, cs, row, okey, acc);
        else select_row(P, okey, ~0ull);
        return;
    }
    if (P.join_type >= CQG_JOIN_RIGHT && !(*(volatile uint32_t*)&s->tag & kJoinMatched)) atomicOr(&s->tag, kJoinMatched);
    uint32_t head = s->head;
    for (uint32_t it = head; it != 0u; it = P.jrow_next[it - 1]) {
        uint64_t roff = P.jrow_off[it - 1];
        // rank of this match among the left row's matches = position in right-file order
        uint32_t rank = 0;
        if (P.jrow_next[head - 1] != 0u) {
            for (uint32_t jt = head; jt != 0u; jt = P.jrow_next[jt - 1]) rank += P.jrow_off[jt - 1] < roff;
        }
        if (rank > 0xffffu) {
            acc.err |= KERR_JOIN_FANOUT;
            rank = 0xffffu;
        }
        if (P.need_right_fields) {
            split_right_row(P, roff, foff, flen);
            rv.rbase = P.rdata + roff;
        }
        bool pass = P.pred_kind == 0 ? true : (P.pred_kind == 1 ? eval_fused(P, row, acc.err) : eval_pred(P.pred, rv, acc.err));
        if (!pass) continue;
        if (P.mode == SCAN_AGG) agg_row(P, cs, row, okey | rank, acc);
        else select_row(P, okey | rank, roff);
    }
}

// RIGHT / FULL joins (evaluator_joins.c:142-171): behind the left-major rows, every right row whose key no left row
// matched, in right-file order, joined to NULL left columns. A key's right rows hang off one slot of the join table;
// the probe marked the slots it matched.
__device__ __noinline__ void emit_right_only_row(const DevPlan& P, const CtaState& cs, uint64_t roff, ThreadAcc& acc) {
    uint32_t foff[2 * kMaxSlots], flen[2 * kMaxSlots];
    for (int k = 0; k < P.nwantL; k++) {
        foff[k] = 0;
        flen[k] = 0;
    }
    split_right_row(P, roff, foff, flen);
    SlowRow row;
    RowView& rv = row.rv;
    rv.base = P.data;
    rv.rbase = P.rdata + roff;
    rv.lfile = P.data;
    rv.rfile = P.rdata;
    rv.foff = foff;
    rv.flen = flen;
    rv.colslot = P.colslot;
    rv.ncols_total = P.n_cols_total;
    rv.nleft_slots = P.nwantL;
    if (roff >> 45) acc.err |= KERR_OFFSET_RANGE;
    const uint64_t okey = kOkeyRightOnly | (roff << 16);
    bool pass = P.pred_kind == 0 ? true : (P.pred_kind == 1 ? eval_fused(P, row, acc.err) : eval_pred(P.pred, rv, acc.err));
    if (!pass) return;
    if (P.mode == SCAN_AGG) agg_row(P, cs, row, okey, acc);
    else select_row(P, okey, roff);
}

// a row that does not fit the staged window: found and split straight from HBM
__device__ __noinline__ void process_long_row(const DevPlan& P, const CtaState& cs, uint64_t goff, ThreadAcc& acc) {
    uint32_t foff[2 * kMaxSlots], flen[2 * kMaxSlots];
    const uint8_t* b = P.data + goff;
    uint64_t maxlen = P.size - goff;
    uint64_t re = 0;
    while (re < maxlen && b[re] != '\n' && b[re] != '\r') re++;
    if (re > 0x7fffffffull) re = 0x7fffffffull;
    split_row_exact(b, 0, (uint32_t)re, P.delim, P.quote, P.wantL, P.nwantL, foff, flen);
    process_row_slow(P, cs, b, P.data, foff, flen, goff, acc);
}

// a row inside the staged window that needs the exact splitter or the general operators
__device__ __noinline__ void process_window_row_slow(const DevPlan& P, const CtaState& cs, const uint8_t* buf, long long g0,
                                                     uint32_t rs, uint32_t len, ThreadAcc& acc) {
    uint32_t foff[2 * kMaxSlots], flen[2 * kMaxSlots];
    split_row_exact(buf, rs, rs + len, P.delim, P.quote, P.wantL, P.nwantL, foff, flen);
    process_row_slow(P, cs, buf, buf - g0, foff, flen, (uint64_t)(g0 + (long long)rs), acc);
}

// parse_line (src/csv_reader.c:285-338) on the row's delimiter and quote bitmasks instead of byte by
// byte: a field is quoted iff its first non-blank byte is the quote character; inside quotes `""` is
// skipped (and counts 2 towards the length of an unterminated field), the first lone quote closes, the
// rest up to the delimiter is dropped. dw/qw: delimiter/quote bits of the row (bit i = byte rs+i), len <= 63.
template <int NW>
__device__ __forceinline__ void split_quoted_masks(const DevPlan& P, const uint8_t* buf, uint32_t rs, uint32_t len,
                                                   unsigned long long dw, unsigned long long qw, uint32_t* off, uint32_t* flen) {
#pragma unroll
    for (int k = 0; k < NW; k++) {
        off[k] = rs;
        flen[k] = 0;
    }
    uint32_t pos = 0;
    int field = 0, wi = 0;
    while (pos < len && wi < P.nwantL) {
        while (pos < len && is_space(buf[rs + pos])) pos++;
        if (pos >= len) break;
        uint32_t fs = pos, fl = 0;
        if ((qw >> pos) & 1ull) {
            pos++;
            fs = pos;
            for (;;) {
                const unsigned long long q = pos < 64u ? (qw & (~0ull << pos)) : 0ull;
                if (!q) {  // unterminated: swallow the rest of the line
                    pos = len;
                    break;
                }
                const uint32_t p = (uint32_t)__ffsll((long long)q) - 1u;
                if (p + 1u < len && ((qw >> (p + 1u)) & 1ull)) {
                    fl += 2;
                    pos = p + 2u;
                    continue;
                }
                fl = p - fs;
                pos = p + 1u;
                break;
            }
            const unsigned long long d = pos < 64u ? (dw & (~0ull << pos)) : 0ull;
            pos = d ? (uint32_t)__ffsll((long long)d) - 1u : len;
        } else {
            const unsigned long long d = dw & (~0ull << pos);
            pos = d ? (uint32_t)__ffsll((long long)d) - 1u : len;
            fl = pos - fs;
        }
        if (P.wantL[wi] == field) {
            if (NW <= 4) {
#pragma unroll
                for (int k = 0; k < NW; k++)
                    if (k == wi) {
                        off[k] = rs + fs;
                        flen[k] = fl;
                    }
            } else {
                off[wi] = rs + fs;
                flen[wi] = fl;
            }
            wi++;
        }
        field++;
        if (pos < len) pos++;  // the delimiter
    }
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
template <class G, int NW>
__global__ void __launch_bounds__(G::THREADS, 4) scan_kernel(const __grid_constant__ DevPlan P) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t* tm = (uint32_t*)(smem + G::OFF_TM);
    uint32_t* dm = (uint32_t*)(smem + G::OFF_DM);
    uint32_t* qm = (uint32_t*)(smem + G::OFF_QM);
    uint16_t* rowpos = (uint16_t*)(smem + G::OFF_ROW);
    uint32_t* wsum = (uint32_t*)(smem + G::OFF_WSUM);
    uint64_t* mbar = (uint64_t*)(smem + G::OFF_MBAR);

    CtaState cs;
    cs.stab = nullptr;
    cs.s_occ = (unsigned long long*)(wsum + 48);
    const bool agg_mode = P.mode == SCAN_AGG;
    if (agg_mode && (P.smem_cap > 0)) cs.stab = smem + G::OFF_TABLE;
    // rows of this plan can take the register route: single table, no generic-interpreter predicate
    const bool fast_plan = !P.join && P.pred_kind != 2 && (P.mode == SCAN_AGG || P.mode == SCAN_SELECT) && P.nwantL <= NW &&
                           !P.exact_only;

    // ---- prologue: barriers, mask padding, shared table ----
    if (tid == 0) {
        for (int s = 0; s < G::STAGES; s++) mbar_init(&mbar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        *cs.s_occ = 0ull;
    }
    for (int w = G::BUF / 32 + tid; w < G::MASKW; w += G::THREADS) {
        tm[w] = 0xffffffffu;
        dm[w] = 0u;
        qm[w] = 0u;
    }
    if (cs.stab) {
        const int words = P.smem_cap * P.entry_bytes / 8;
        const int ew = P.entry_bytes / 8;
        for (int k = tid; k < words; k += G::THREADS) {
            int o = k % ew;
            ((uint64_t*)cs.stab)[k] = o == 0 ? 0ull : ((const uint64_t*)P.entry_init)[o];
        }
    }
    __syncthreads();

    ThreadAcc acc;
    acc.rows = 0;
    acc.count = 0;
    acc.first = ~0ull;
    acc.err = 0;
#pragma unroll
    for (int a = 0; a < 4; a++) {
        acc.si[a] = 0;
        acc.sd[a] = 0.0;
        acc.s3[a] = 0;
        acc.sn[a] = 0;
    }

    const uint64_t size = P.size;
    const uint64_t copy_end = (size + 15ull) & ~15ull;  // allocation is padded (cqg_device_padding)
    const uint32_t patD = (uint32_t)P.delim * 0x01010101u;
    const uint32_t patQ = (uint32_t)P.quote * 0x01010101u;

    auto issue = [&](int it) {
        const long long ti = (long long)blockIdx.x + (long long)it * gridDim.x;
        long long tile = P.tile_list ? (long long)P.tile_list[ti] : (long long)P.first_tile + ti;
        int stage = it % G::STAGES;
        long long g0 = tile * (long long)G::TILE - G::PRE;  // file offset of buf[0]
        uint32_t skip = g0 < 0 ? (uint32_t)(-g0) : 0u;
        long long src = g0 + skip;
        long long avail = (long long)copy_end - src;
        uint32_t bytes = 0;
        if (avail > 0) bytes = (uint32_t)(avail < (long long)(G::BUF - skip) ? avail : (long long)(G::BUF - skip));
        uint8_t* dst = smem + G::OFF_BUF + stage * G::BUF + skip;
        if (bytes) {
            mbar_expect_tx(&mbar[stage], bytes);
            tma_load_1d(dst, P.data + src, bytes, &mbar[stage]);
        } else {
            mbar_expect_tx(&mbar[stage], 0);
        }
    };

    const int my_tiles = (P.n_tiles > (int)blockIdx.x) ? (P.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (tid == 0)
        for (int it = 0; it < G::STAGES - 1 && it < my_tiles; it++) issue(it);

    for (int it = 0; it < my_tiles; it++) {
        const int stage = it % G::STAGES;
        const uint32_t parity = (uint32_t)(it / G::STAGES) & 1u;
        if (tid == 0 && it + G::STAGES - 1 < my_tiles) issue(it + G::STAGES - 1);
        // stop early when a table overflowed: the host will retry with a bigger one
        const unsigned abort_now = (*(volatile unsigned*)P.errflags) & (KERR_TABLE_FULL | KERR_SEL_OVERFLOW);
        mbar_wait(&mbar[stage], parity);
        const uint8_t* buf = smem + G::OFF_BUF + stage * G::BUF;
        const long long ti = (long long)blockIdx.x + (long long)it * gridDim.x;
        const long long tile = P.tile_list ? (long long)P.tile_list[ti] : (long long)P.first_tile + ti;
        const long long g0 = tile * (long long)G::TILE - G::PRE;
        const bool edge_tile = g0 < 0 || g0 + G::BUF > (long long)size;  // some staged bytes lie outside the file

        // ---------------- phase 1: classify ----------------
        uint32_t spec = 0;
        if (!edge_tile && !P.exact_only) {
#pragma unroll 2
            for (int c = tid; c < G::CHUNKS; c += G::THREADS) {
                const uint4 v = *(const uint4*)(buf + 16 * c);
                uint32_t f0 = eq_flags7(v.x, 0x0a0a0a0au), f1 = eq_flags7(v.y, 0x0a0a0a0au), f2 = eq_flags7(v.z, 0x0a0a0a0au),
                         f3 = eq_flags7(v.w, 0x0a0a0a0au);
                uint32_t T = flags_to_mask16(f0, f1, f2, f3);
                uint32_t D = flags_to_mask16(eq_flags7(v.x, patD), eq_flags7(v.y, patD), eq_flags7(v.z, patD), eq_flags7(v.w, patD));
                // any byte < 0x23 other than '\n' (CR, quote, blanks, NUL, controls): the tile takes the careful path
                uint32_t x0 = v.x | f0, x1 = v.y | f1, x2 = v.z | f2, x3 = v.w | f3;
                spec |= ((x0 - 0x23232323u) & ~x0) | ((x1 - 0x23232323u) & ~x1) | ((x2 - 0x23232323u) & ~x2) |
                        ((x3 - 0x23232323u) & ~x3);
                ((uint16_t*)tm)[c] = (uint16_t)T;
                ((uint16_t*)dm)[c] = (uint16_t)D;
            }
            spec &= 0x80808080u;
        } else {
            spec = 1;  // edge tiles and exotic dialects: always the careful path
            for (int c = tid; c < G::CHUNKS; c += G::THREADS) {
                const uint4 v = *(const uint4*)(buf + 16 * c);
                uint32_t T = eq_mask16(v, 0x0a0a0a0au), D = eq_mask16(v, patD);
                long long g = g0 + 16 * c;
                if (g < 0 || g + 16 > (long long)size) {
                    // bytes outside the file are row terminators and nothing else
                    uint32_t valid = 0;
                    for (int j = 0; j < 16; j++)
                        if (g + j >= 0 && g + j < (long long)size) valid |= 1u << j;
                    T = (T & valid) | (~valid & 0xffffu);
                    D &= valid;
                }
                ((uint16_t*)tm)[c] = (uint16_t)T;
                ((uint16_t*)dm)[c] = (uint16_t)D;
            }
        }
        const int special = __syncthreads_or((int)(spec != 0u));
        if (special) {
            for (int c = tid; c < G::CHUNKS; c += G::THREADS) {
                const uint4 v = *(const uint4*)(buf + 16 * c);
                uint32_t R = eq_mask16(v, 0x0d0d0d0du), Q = eq_mask16(v, patQ);
                long long g = g0 + 16 * c;
                if (g < 0 || g + 16 > (long long)size) {
                    uint32_t valid = 0;
                    for (int j = 0; j < 16; j++)
                        if (g + j >= 0 && g + j < (long long)size) valid |= 1u << j;
                    R &= valid;
                    Q &= valid;
                }
                ((uint16_t*)tm)[c] |= (uint16_t)R;
                ((uint16_t*)qm)[c] = (uint16_t)Q;
            }
            __syncthreads();
        }

        // ---------------- phase 1b: row starts ----------------
        long long olo_l = (long long)P.own_lo - g0, ohi_l = (long long)P.own_hi - g0;
        const uint32_t olo = (uint32_t)(olo_l < G::PRE ? G::PRE : (olo_l > G::PRE + G::TILE ? G::PRE + G::TILE : olo_l));
        const uint32_t ohi = (uint32_t)(ohi_l < G::PRE ? G::PRE : (ohi_l > G::PRE + G::TILE ? G::PRE + G::TILE : ohi_l));
        uint32_t S[G::WPT];
        uint32_t mycount = 0;
        {
            const int w0 = G::PRE / 32 + tid * G::WPT;
            uint32_t prev = tm[w0 - 1];
#pragma unroll
            for (int j = 0; j < G::WPT; j++) {
                uint32_t t = tm[w0 + j];
                uint32_t s = ((t << 1) | (prev >> 31)) & ~t;
                prev = t;
                uint32_t p0 = (uint32_t)(w0 + j) * 32u;
                // keep bits in [olo, ohi)
                uint32_t lo_m = olo <= p0 ? 0xffffffffu : (olo >= p0 + 32u ? 0u : (0xffffffffu << (olo - p0)));
                uint32_t hi_m = ohi >= p0 + 32u ? 0xffffffffu : (ohi <= p0 ? 0u : (0xffffffffu >> (p0 + 32u - ohi)));
                s &= lo_m & hi_m;
                S[j] = s;
                mycount += __popc(s);
            }
        }
        // block exclusive scan of mycount
        uint32_t incl = mycount;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t n = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += n;
        }
        if (lane == 31) wsum[warp] = incl;
        const int abort_all = __syncthreads_or((int)abort_now);
        uint32_t wbase = 0, nrows = 0;
#pragma unroll
        for (int w = 0; w < G::NWARPS; w++) {
            uint32_t x = wsum[w];
            if (w < warp) wbase += x;
            nrows += x;
        }
        const uint32_t mybase = wbase + incl - mycount;
        if (abort_all) nrows = 0;

        for (uint32_t pass_lo = 0; pass_lo < nrows; pass_lo += G::ROWCAP) {
            {
                uint32_t idx = mybase;
                const int w0 = G::PRE / 32 + tid * G::WPT;
#pragma unroll
                for (int j = 0; j < G::WPT; j++) {
                    uint32_t s = S[j];
                    while (s) {
                        uint32_t b = __ffs(s) - 1;
                        s &= s - 1;
                        uint32_t k = idx - pass_lo;
                        if (k < (uint32_t)G::ROWCAP) rowpos[k] = (uint16_t)((w0 + j) * 32 + b);
                        idx++;
                    }
                }
            }
            __syncthreads();
            // ---------------- phase 2: rows ----------------
            const uint32_t npass = nrows - pass_lo < (uint32_t)G::ROWCAP ? nrows - pass_lo : (uint32_t)G::ROWCAP;
            if (P.mode == SCAN_COUNT_ROWS) {
                for (uint32_t r = tid; r < npass; r += G::THREADS) acc.rows++;
            } else {
                for (uint32_t r = tid; r < npass; r += G::THREADS) {
                    const uint32_t rs = rowpos[r];
                    acc.rows++;
                    // row end: first terminator at or after rs
                    const unsigned long long tw = mask_window(tm, rs);
                    if (tw == 0ull) {
                        // longer than 64 bytes: look further, then take the general route
                        uint32_t p = rs + 64u, len = 0xffffffffu;
                        while (p < (uint32_t)G::BUF) {
                            unsigned long long w2 = mask_window(tm, p);
                            if (w2) {
                                len = p + (uint32_t)__ffsll((long long)w2) - 1u - rs;
                                break;
                            }
                            p += 64u;
                        }
                        if (len == 0xffffffffu || rs + len >= (uint32_t)G::BUF) process_long_row(P, cs, (uint64_t)(g0 + (long long)rs), acc);
                        else process_window_row_slow(P, cs, buf, g0, rs, len, acc);
                        continue;
                    }
                    const uint32_t len = (uint32_t)__ffsll((long long)tw) - 1u;
                    if (rs + len >= (uint32_t)G::BUF) {
                        process_long_row(P, cs, (uint64_t)(g0 + (long long)rs), acc);
                        continue;
                    }
                    const unsigned long long lenmask = (1ull << len) - 1ull;  // len <= 63 here
                    if (!fast_plan) {
                        process_window_row_slow(P, cs, buf, g0, rs, len, acc);
                        continue;
                    }
                    // ---- register route ----
                    FastRow<NW> row;
                    row.base = buf;
                    row.lfile = buf - g0;
                    row.clean = !special;
                    const unsigned long long qw = special ? (mask_window(qm, rs) & lenmask) : 0ull;
                    if (qw) {
                        // quote characters in the row: the quote-aware split, still on bitmasks
                        split_quoted_masks<NW>(P, buf, rs, len, mask_window(dm, rs) & lenmask, qw, row.off, row.len);
                    } else {
                        // wanted field k starts after gap[k] further delimiters: clear gap-1 of the
                        // remaining delimiter bits, the next one marks the start, the one after the end
                        unsigned long long dw = mask_window(dm, rs) & lenmask;
                        uint32_t startpos = 0;
                        bool missing = false;
#pragma unroll
                        for (int k = 0; k < NW; k++) {
                            row.off[k] = rs;
                            row.len[k] = 0;
                            if (k < P.nwantL) {
                                const int gap = P.gap[k];
                                if (gap > 0) {
                                    for (int i = 1; i < gap; i++) dw &= dw - 1ull;
                                    missing = missing || dw == 0ull;
                                    startpos = (uint32_t)__ffsll((long long)dw);
                                    dw &= dw - 1ull;
                                }
                                uint32_t endpos = dw ? (uint32_t)__ffsll((long long)dw) - 1u : len;
                                uint32_t fs = startpos;
                                if (special)
                                    while (fs < endpos && is_space(buf[rs + fs])) fs++;
                                row.off[k] = rs + fs;
                                row.len[k] = missing ? 0u : endpos - fs;
                            }
                        }
                    }
                    const uint64_t gabs = P.global_base + (uint64_t)(g0 + (long long)rs);
                    if (gabs >> 45) acc.err |= KERR_OFFSET_RANGE;
                    const uint64_t okey = gabs << 16;
                    if (P.pred_kind == 1 && !eval_fused(P, row, acc.err)) continue;
                    if (P.mode == SCAN_AGG) agg_row(P, cs, row, okey, acc);
                    else select_row(P, okey, 0);
                }
            }
            __syncthreads();
        }
        __syncthreads();
    }

    // ---------------- epilogue ----------------
    // rows seen
    {
        uint32_t r = acc.rows;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) r += __shfl_xor_sync(0xffffffffu, r, d);
        if (lane == 0 && r) atomicAdd(P.rows_scanned, (unsigned long long)r);
    }
    if (agg_mode && P.scalar_regs) {
        // fold the register accumulators into the single shared entry
        uint8_t* e = cs.stab;
        uint32_t c = acc.count;
        uint64_t f = acc.first;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            c += __shfl_xor_sync(0xffffffffu, c, d);
            uint64_t of = __shfl_xor_sync(0xffffffffu, f, d);
            f = of < f ? of : f;
        }
        if (lane == 0 && c) {
            atomicAdd((unsigned int*)(e + kOffCount), c);
            amin64((uint64_t*)(e + kOffFirst), f);
        }
#pragma unroll
        for (int a = 0; a < 4; a++) {
            if (a < P.naggs && P.aggs[a].off >= 0 && (P.aggs[a].func == CQG_AGG_SUM || P.aggs[a].func == CQG_AGG_AVG)) {
                long long si = acc.si[a], s3 = acc.s3[a];
                double sd = acc.sd[a];
                uint32_t sn = acc.sn[a];
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    si += __shfl_xor_sync(0xffffffffu, si, d);
                    s3 += __shfl_xor_sync(0xffffffffu, s3, d);
                    sd += __shfl_xor_sync(0xffffffffu, sd, d);
                    sn += __shfl_xor_sync(0xffffffffu, sn, d);
                }
                if (lane == 0 && sn) {
                    uint8_t* st = e + P.aggs[a].off;
                    atomicAdd((unsigned long long*)st, (unsigned long long)si);
                    atomicAdd((double*)(st + 8), sd);
                    atomicAdd((unsigned int*)(st + 16), sn);
                    atomicAdd((unsigned long long*)(st + 24), (unsigned long long)s3);  // short decimals, x1000
                }
            }
        }
        __syncthreads();
        // the single group `_all_`
        if (tid == 0) {
            uint64_t h = key_hash_final(0x243F6A8885A308D3ull);
            uint8_t* ge = global_entry_for(P, h, 0u, nullptr, acc.err);
            if (ge) entry_merge(P, ge, e);
        }
    } else if (cs.stab) {
        __syncthreads();
        for (int s = tid; s < P.smem_cap; s += G::THREADS) {
            const uint8_t* e = cs.stab + (size_t)s * P.entry_bytes;
            uint64_t h = *(const uint64_t*)(e + kOffHash);
            if (h == 0ull) continue;
            uint64_t kw[2 * CQG_MAX_GROUP_COLS];
            for (int g = 0; g < P.ngc; g++) {
                kw[2 * g] = *(const uint64_t*)(e + kOffKeys + 16 * g);
                kw[2 * g + 1] = *(const uint64_t*)(e + kOffKeys + 16 * g + 8);
            }
            uint8_t* ge = global_entry_for(P, h, *(const uint32_t*)(e + kOffTags), kw, acc.err);
            if (ge) entry_merge(P, ge, e);
        }
    }
    if (acc.err) atomicOr(P.errflags, acc.err);
}

#ifndef CQG_JIT  // a kernel compiled at run time for one query (cqg_jit) needs none of the kernels below
// ------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------
__global__ void init_table_kernel(uint8_t* tab, uint64_t cap, int entry_bytes, const uint8_t* init) {
    const uint64_t words = cap * (uint64_t)(entry_bytes / 8);
    const int ew = entry_bytes / 8;
    for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < words; k += (uint64_t)gridDim.x * blockDim.x)
        ((uint64_t*)tab)[k] = ((const uint64_t*)init)[k % ew];
}

// compact the occupied entries of a table into `out` (entry images); *n_out counts them.
// owner filter: (hash % world) == owner when world > 1.
__global__ void compact_table_kernel(const uint8_t* tab, uint64_t cap, int entry_bytes, uint8_t* out, uint64_t out_cap,
                                     unsigned long long* n_out, int owner, int world) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint8_t* e = tab + i * (uint64_t)entry_bytes;
        uint64_t h = *(const uint64_t*)e;
        if (h == 0ull) continue;
        if (world > 1 && (int)((h >> 8) % (uint64_t)world) != owner) continue;
        unsigned long long idx = atomicAdd(n_out, 1ull);
        if (idx < out_cap) {
            uint8_t* o = out + idx * (uint64_t)entry_bytes;
            for (int k = 0; k < entry_bytes; k += 8) *(uint64_t*)(o + k) = *(const uint64_t*)(e + k);
        }
    }
}

__global__ void owner_count_kernel(const uint8_t* tab, uint64_t cap, int entry_bytes, int world, unsigned long long* counts) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t h = *(const uint64_t*)(tab + i * (uint64_t)entry_bytes);
        if (h == 0ull) continue;
        atomicAdd(&counts[world > 1 ? (h >> 8) % (uint64_t)world : 0], 1ull);
    }
}

// first-appearance keys of dense entries, and the gather that puts entries in sorted order
__global__ void extract_first_kernel(const uint8_t* ent, uint64_t n, int entry_bytes, uint64_t* keys, uint32_t* idx) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        keys[i] = *(const uint64_t*)(ent + i * (uint64_t)entry_bytes + kOffFirst);
        idx[i] = (uint32_t)i;
    }
}
__global__ void gather_entries_kernel(const uint8_t* ent, const uint32_t* idx, uint64_t n, int entry_bytes, uint8_t* out) {
    const uint64_t per = (uint64_t)entry_bytes / 16;
    const uint64_t total = n * per;
    for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < total; k += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t e = k / per, w = k % per;
        ((uint4*)out)[k] = ((const uint4*)(ent + (uint64_t)idx[e] * entry_bytes))[w];
    }
}

// fold n serialized entries into the table of plan P
__global__ void merge_entries_kernel(const __grid_constant__ DevPlan P, const uint8_t* recs, uint64_t n) {
    unsigned err = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint8_t* e = recs + i * (uint64_t)P.entry_bytes;
        uint64_t h = *(const uint64_t*)(e + kOffHash);
        if (h == 0ull) continue;  // an unused record of a fixed-capacity exchange buffer (a key's hash is never 0)
        uint64_t kw[2 * CQG_MAX_GROUP_COLS];
        for (int g = 0; g < P.ngc; g++) {
            kw[2 * g] = *(const uint64_t*)(e + kOffKeys + 16 * g);
            kw[2 * g + 1] = *(const uint64_t*)(e + kOffKeys + 16 * g + 8);
        }
        uint8_t* ge = global_entry_for(P, h, *(const uint32_t*)(e + kOffTags), kw, err);
        if (ge) entry_merge(P, ge, e);
    }
    if (err) atomicOr(P.errflags, err);
}

// ------------------------------------------------------------------------------------------
// fetch: decode columns of rows given by file offset (group first rows, selected rows)
// ------------------------------------------------------------------------------------------
struct OutCell {
    int32_t type;
    uint32_t len;      // STRING: trimmed length
    uint64_t payload;  // INTEGER / DOUBLE bits / DATE packed / STRING: table bit 63 | file offset
};

struct FetchParams {
    const uint8_t* data;
    uint64_t size;
    const uint8_t* rdata;
    uint64_t rsize;
    uint8_t delim, quote;
    int32_t n_left_cols;
    int32_t ncols;                    // output columns
    int16_t cols[CQG_MAX_OUT_COLS];   // query column index (-1: NULL)
    const uint64_t* loff;             // [n] left row offsets
    const uint64_t* roff;             // [n] right row offsets or nullptr
    uint64_t n;
    OutCell* out;                     // [n][ncols]
    unsigned* errflags;
};

__device__ inline void fetch_one(const uint8_t* file, uint64_t fsize, uint64_t off, uint8_t delim, uint8_t quote, int col,
                                 bool right, OutCell& oc, unsigned& err) {
    oc.type = T_NULL;
    oc.len = 0;
    oc.payload = 0;
    if (col < 0 || off >= fsize) return;
    const uint8_t* b = file + off;
    uint64_t maxlen = fsize - off;
    uint64_t re = 0;
    while (re < maxlen && b[re] != '\n' && b[re] != '\r') re++;
    if (re > 0x7fffffffull) re = 0x7fffffffull;
    int16_t want = (int16_t)col;
    uint32_t fo, fl;
    split_row_exact(b, 0, (uint32_t)re, delim, quote, &want, 1, &fo, &fl);
    DVal v = decode_field(b + fo, fl, err);
    oc.type = v.type;
    if (v.type == T_STR) {
        oc.len = v.len;
        oc.payload = (right ? (1ull << 63) : 0ull) | (uint64_t)(v.s - file);
    } else if (v.type != T_NULL) {
        oc.payload = (uint64_t)v.i;
    }
}

// Fields per row as parse_line counts them (src/csv_reader.c:278-338: leading whitespace skipped, quoted fields with doubled
// quotes, a trailing delimiter opens no field, a line of blanks has none) for rows given by their byte offsets: csv_load's
// Row::column_count (src/csv_reader.c:358-366), which a GPU-backed csv_load must hand to its callers.
__global__ void field_count_kernel(const uint8_t* data, uint64_t size, const uint64_t* row_off, uint64_t n, uint8_t delim, uint8_t quote,
                                   int32_t* out) {
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t ptr = row_off[r];
        uint64_t re = ptr;
        while (re < size && data[re] != '\n' && data[re] != '\r') re++;
        int fc = 0;
        while (ptr < re) {
            while (ptr < re && is_space(data[ptr])) ptr++;
            if (ptr >= re) break;
            if (data[ptr] == quote) {
                ptr++;
                while (ptr < re) {
                    if (data[ptr] == quote) {
                        if (ptr + 1 < re && data[ptr + 1] == quote) {
                            ptr += 2;
                        } else {
                            ptr++;
                            break;
                        }
                    } else {
                        ptr++;
                    }
                }
            }
            while (ptr < re && data[ptr] != delim) ptr++;
            fc++;
            if (ptr < re && data[ptr] == delim) ptr++;
        }
        out[r] = fc;
    }
}


__global__ void fetch_kernel(const __grid_constant__ FetchParams F) {
    unsigned err = 0;
    const uint64_t total = F.n * (uint64_t)F.ncols;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t r = i / (uint64_t)F.ncols;
        int c = (int)(i % (uint64_t)F.ncols);
        int col = F.cols[c];
        OutCell oc;
        if (col >= F.n_left_cols) {
            if (F.roff && F.rdata) fetch_one(F.rdata, F.rsize, F.roff[r], F.delim, F.quote, col - F.n_left_cols, true, oc, err);
            else {
                oc.type = T_NULL;
                oc.len = 0;
                oc.payload = 0;
            }
        } else {
            fetch_one(F.data, F.size, F.loff[r], F.delim, F.quote, col, false, oc, err);
        }
        F.out[i] = oc;
    }
    if (err) atomicOr(F.errflags, err);
}

// ------------------------------------------------------------------------------------------
// finish on the device (results of many groups, build_aggregated_result evaluator_aggregates.c:533-696): the
// result arrays of cqg_result_t are written in their final layout into one device block that goes to the host
// in one piece. `idx` orders the dense entries by first appearance.
// ------------------------------------------------------------------------------------------
struct FinishParams {
    const uint8_t* entries;  // dense general entries
    const uint32_t* idx;     // [G] entry of output row gi
    uint64_t G;
    int32_t entry_bytes;
    int32_t n_aggs, n_out;
    int32_t agg_func[CQG_MAX_AGGS], agg_col[CQG_MAX_AGGS], agg_off[CQG_MAX_AGGS];
    int16_t out_cols[CQG_MAX_OUT_COLS];
    const uint8_t* data;     // the (left) file
    uint64_t size;
    uint64_t global_base;
    uint8_t delim, quote;
    // outputs (device block laid out like the host block)
    uint64_t* first_offset;  // [G]
    int64_t* count;          // [G]
    double* sum;             // [n_aggs][G]
    int64_t* ncount;         // [n_aggs][G]
    OutCell* cells;          // [n_aggs + n_out][G]
    uint64_t* str_size;      // [n_aggs + n_out][G]: bytes the cell's string takes (len + 1; 0: not a string)
    unsigned* errflags;
};

__global__ void finish_cells_kernel(const __grid_constant__ FinishParams F) {
    unsigned err = 0;
    const uint64_t total = F.G * (uint64_t)(F.n_aggs + F.n_out);
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const int cell = (int)(i / F.G);
        const uint64_t gi = i % F.G;
        const uint8_t* e = F.entries + (uint64_t)F.idx[gi] * (uint64_t)F.entry_bytes;
        const uint64_t first = *(const uint64_t*)(e + kOffFirst);
        const long long count = (long long)*(const uint64_t*)(e + kOffCount);
        OutCell oc;
        oc.type = T_NULL;
        oc.len = 0;
        oc.payload = 0;
        if (cell == 0) {
            F.first_offset[gi] = first == ~0ull ? 0ull : (first >> 16);
            F.count[gi] = count;
        }
        if (cell < F.n_aggs) {
            const int f = F.agg_func[cell], off = F.agg_off[cell];
            double sum = 0.0;
            long long n = 0;
            if (f == CQG_AGG_COUNT_STAR || (f == CQG_AGG_COUNT && F.agg_col[cell] >= 0)) {
                oc.type = T_INT;
                oc.payload = (uint64_t)(long long)(int)count;  // `result.int_value = row_count` with an int row_count (aggregates.c:270)
            } else if (F.agg_col[cell] < 0) {
                // unknown column: NULL
            } else if (f == CQG_AGG_SUM || f == CQG_AGG_AVG) {
                const long long si = *(const long long*)(e + off);
                const double sd = *(const double*)(e + off + 8);
                n = (long long)*(const uint64_t*)(e + off + 16);
                const long long s3 = *(const long long*)(e + off + 24);
                // the same three IEEE operations as the host path, no contraction
                sum = __dadd_rn(__dadd_rn(__ll2double_rn(si), sd), __ddiv_rn(__ll2double_rn(s3), 1000.0));
                const double v = f == CQG_AGG_SUM ? sum : ((int)n > 0 ? __ddiv_rn(sum, (double)(int)n) : 0.0);
                oc.type = T_DBL;
                oc.payload = (uint64_t)__double_as_longlong(v);
            } else {
                const uint64_t* st = (const uint64_t*)(e + off);
                const uint32_t cls = st[0] == ~0ull ? 0u : (uint32_t)(st[0] & 3u);
                if (cls == 1u) {
                    // type and bits of the extreme come from the row that holds it
                    fetch_one(F.data, F.size, (st[3] >> 16) - F.global_base, F.delim, F.quote, F.agg_col[cell], false, oc, err);
                } else if (cls == 3u) {
                    oc.type = T_DATE;
                    oc.payload = st[1] - 1ull;
                } else if (cls == 2u) {
                    oc.type = T_STR;
                    oc.len = (uint32_t)(st[4] & 0x3ffffu);
                    oc.payload = (st[4] & (1ull << 63)) | ((st[4] >> 18) & 0x1fffffffffffull);
                }
            }
            F.sum[i] = sum;
            F.ncount[i] = n;
        } else if (first != ~0ull) {
            fetch_one(F.data, F.size, (first >> 16) - F.global_base, F.delim, F.quote, F.out_cols[cell - F.n_aggs], false, oc, err);
        }
        F.cells[i] = oc;
        F.str_size[i] = oc.type == T_STR ? (uint64_t)oc.len + 1ull : 0ull;
    }
    if (err) atomicOr(F.errflags, err);
}

// cells -> cqg_value_t (24 bytes: type, reserved, 16-byte union) in the device copy of the host block; strings are
// copied behind the arrays, NUL-terminated, and pointed to by the address they will have ON THE HOST.
__global__ void finish_values_kernel(const OutCell* cells, const uint64_t* str_off, uint64_t n, const uint8_t* data, uint8_t* values,
                                     uint8_t* strings, uint64_t host_strings) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const OutCell oc = cells[i];
        uint64_t w0 = 0, w1 = 0;
        int32_t type = oc.type;
        switch (oc.type) {
            case T_INT:
            case T_DBL: w0 = oc.payload; break;
            case T_DATE:
                w0 = ((oc.payload >> 16) & 0xffffffffull) | (((oc.payload >> 8) & 0xffull) << 32);  // year | month
                w1 = oc.payload & 0xffull;                                                          // day
                break;
            case T_STR: {
                const uint64_t so = str_off[i];
                const uint8_t* src = data + (oc.payload & 0x7fffffffffffffffull);
                uint8_t* d = strings + so;
                for (uint32_t k = 0; k < oc.len; k++) d[k] = src[k];
                d[oc.len] = 0;
                w0 = host_strings + so;
                break;
            }
            default: type = T_NULL; break;
        }
        uint8_t* v = values + 24ull * i;
        *(uint64_t*)v = (uint64_t)(uint32_t)type;
        *(uint64_t*)(v + 8) = w0;
        *(uint64_t*)(v + 16) = w1;
    }
}

// for aggregated joins: right row of the group's first joined row = the rank-th match of the left row
__global__ void resolve_first_right_kernel(const __grid_constant__ DevPlan P, const uint64_t* first_okey, uint64_t n,
                                           uint64_t* roff_out) {
    unsigned err = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t ok = first_okey[i];
        roff_out[i] = ~0ull;
        if (ok == ~0ull) continue;
        if (ok & kOkeyRightOnly) {  // a right row without a match: the okey carries its offset
            roff_out[i] = (ok >> 16) & ((1ull << 45) - 1ull);
            continue;
        }
        uint64_t loff = (ok >> 16) - P.global_base;
        uint32_t rank = (uint32_t)(ok & 0xffffu);
        const uint8_t* b = P.data + loff;
        uint64_t maxlen = P.size - loff;
        uint64_t re = 0;
        while (re < maxlen && b[re] != '\n' && b[re] != '\r') re++;
        int16_t want = (int16_t)P.jl_col;
        uint32_t fo, fl;
        split_row_exact(b, 0, (uint32_t)re, P.delim, P.quote, &want, 1, &fo, &fl);
        DVal lk = decode_field(b + fo, fl, err);
        uint32_t tag;
        uint64_t w0, w1, h;
        join_key_of(lk, err, tag, w0, w1, h);
        JoinSlot* s = join_find(P, h, tag, w0, w1, false);
        if (!s) continue;
        for (uint32_t it = s->head; it != 0u; it = P.jrow_next[it - 1]) {
            uint64_t ro = P.jrow_off[it - 1];
            uint32_t rk = 0;
            for (uint32_t jt = s->head; jt != 0u; jt = P.jrow_next[jt - 1]) rk += P.jrow_off[jt - 1] < ro;
            if (rk == rank) {
                roff_out[i] = ro;
                break;
            }
        }
    }
    if (err) atomicOr(P.errflags, err);
}

// gather string bytes: dst[dst_off[i] .. +len) = file bytes. One thread per string (result strings are short).
__global__ void pack_strings_kernel(const uint8_t* data, const uint8_t* rdata, const uint64_t* refs, const uint32_t* lens,
                                    const uint64_t* dst_off, uint64_t n, uint8_t* dst) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t ref = refs[i];
        const uint8_t* src = ((ref >> 63) ? rdata : data) + (ref & 0x7fffffffffffffffull);
        uint8_t* d = dst + dst_off[i];
        const uint32_t len = lens[i];
        for (uint32_t k = 0; k < len; k++) d[k] = src[k];
    }
}

// parse_value on a single field (cqg_parse_value)
__global__ void parse_value_kernel(const uint8_t* s, uint32_t len, OutCell* out, unsigned* errflags) {
    unsigned err = 0;
    DVal v = decode_field(s, len, err);
    out->type = v.type;
    out->len = v.type == T_STR ? v.len : 0;
    out->payload = v.type == T_STR ? (uint64_t)(v.s - s) : (uint64_t)v.i;
    if (err) atomicOr(errflags, err);
}

// ------------------------------------------------------------------------------------------
// synthetic data: restatement of utils/generate_big_dataset.py:9-19, row i from (seed, i) only
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t gen_mix(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__host__ __device__ inline uint32_t gen_row(uint8_t* o, uint64_t seed, uint64_t i, long long key_card) {
    uint64_t r = gen_mix(seed * 0xD1342543DE82EF95ull + i);
    uint64_t r2 = gen_mix(r);
    uint8_t* p = o;
    uint8_t nm = (uint8_t)('A' + (r & 15));
    uint8_t sn = (uint8_t)('A' + ((r >> 4) & 15));
    unsigned age = 10 + (unsigned)(((r >> 8) & 0xffffff) % 71);
    uint8_t gd = ((r >> 32) & 1) ? 'm' : 'f';
    unsigned h = 100 + (unsigned)(((r >> 33) & 0xffffff) % 101);
    for (int k = 0; k < 10; k++) *p++ = nm;
    *p++ = ',';
    for (int k = 0; k < 8; k++) *p++ = sn;
    *p++ = ',';
    *p++ = (uint8_t)('0' + age / 10);
    *p++ = (uint8_t)('0' + age % 10);
    *p++ = ',';
    *p++ = gd;
    *p++ = ',';
    *p++ = (uint8_t)('0' + h / 100);
    *p++ = '.';
    unsigned frac = h % 100;
    *p++ = (uint8_t)('0' + frac / 10);
    if (frac % 10) *p++ = (uint8_t)('0' + frac % 10);
    if (key_card > 0) {
        *p++ = ',';
        uint64_t uid = r2 % (uint64_t)key_card;
        uint8_t tmp[24];
        int n = 0;
        do {
            tmp[n++] = (uint8_t)('0' + uid % 10);
            uid /= 10;
        } while (uid);
        while (n) *p++ = tmp[--n];
    }
    *p++ = '\n';
    return (uint32_t)(p - o);
}

constexpr int kGenRowsPerBlock = 1024;
// pass 1: bytes of each block of rows
__global__ void gen_sizes_kernel(long long row_start, long long rows, uint64_t seed, long long key_card, unsigned long long* block_bytes) {
    __shared__ unsigned int sum;
    if (threadIdx.x == 0) sum = 0;
    __syncthreads();
    uint8_t tmp[64];
    unsigned my = 0;
    long long base = (long long)blockIdx.x * kGenRowsPerBlock;
    for (int k = threadIdx.x; k < kGenRowsPerBlock; k += blockDim.x)
        if (base + k < rows) my += gen_row(tmp, seed, (uint64_t)(row_start + base + k), key_card);
    atomicAdd(&sum, my);
    __syncthreads();
    if (threadIdx.x == 0) block_bytes[blockIdx.x] = sum;
}
// pass 2: write the rows of block b at block_off[b]
__global__ void gen_write_kernel(uint8_t* out, long long row_start, long long rows, uint64_t seed, long long key_card,
                                 const unsigned long long* block_off) {
    __shared__ unsigned int lens[kGenRowsPerBlock];
    __shared__ unsigned int offs[kGenRowsPerBlock];
    uint8_t tmp[64];
    long long base = (long long)blockIdx.x * kGenRowsPerBlock;
    for (int k = threadIdx.x; k < kGenRowsPerBlock; k += blockDim.x)
        lens[k] = (base + k < rows) ? gen_row(tmp, seed, (uint64_t)(row_start + base + k), key_card) : 0u;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int o = 0;
        for (int k = 0; k < kGenRowsPerBlock; k++) {
            offs[k] = o;
            o += lens[k];
        }
    }
    __syncthreads();
    uint8_t* dst = out + block_off[blockIdx.x];
    for (int k = threadIdx.x; k < kGenRowsPerBlock; k += blockDim.x) {
        if (base + k < rows) {
            uint32_t n = gen_row(tmp, seed, (uint64_t)(row_start + base + k), key_card);
            for (uint32_t j = 0; j < n; j++) dst[offs[k] + j] = tmp[j];
        }
    }
}

#endif  // CQG_JIT

}  // namespace cqg
