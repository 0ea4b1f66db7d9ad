// cqg_lean2k.cuh — the written-out few-groups GROUP BY loop (BASELINE configs[0] shape:
//   SELECT role, COUNT(*), AVG(age) FROM f WHERE age > 25 GROUP BY role)
// for DevPlan::simple == 2 plans with ONE key column, COUNT / SUM / AVG over <= 3 columns and a WHERE that is
// empty or a program of `column <op> decimal literal` leaves: create_groups (evaluator_aggregates.c:108-176) and
// evaluate_aggregate (:263-326) on lean2g_kernel's tile pipeline (cqg_lean2g.cuh: TMA tile, exact '\n' /
// delimiter / "other special byte" classes, clean-tile verdict before the walk, thread-owned cursor walk), with
// a row loop that costs a third of lean2g's instructions:
//   * the key is the field's RAW bytes (<= 16, zero padded: 5 shared loads, 4 funnel shifts, one mask): no
//     per-row classing, trimming or number parsing. What the reference makes of those bytes (trim, NULL,
//     INTEGER / DOUBLE images: canon_part) is worked out ONCE per group and CTA when the group is numbered, and
//     groups whose raw keys differ but whose canonical keys agree ("7", "07") meet in the global table;
//   * one multiplicative hash (4 IMAD), a dictionary of 128 two-word buckets of `hash tag | group << 2 | fresh | 1`
//     (one 64-bit load gives both words of the home bucket: with <= 16 keys a third key in one bucket is rare, and
//     only such a key takes the out-of-line probe loop) and one 16-byte compare against the group's key record;
//   * per-warp accumulators with native 32-bit shared atomics, laid out [word][group] so that different groups never
//     share a bank: COUNT is one ATOMS without a result, a sum one ATOMS and a carry test (the lane that wraps the
//     low word adds to the high one). 2.3 KB per CTA: the kernel keeps lean2g's six to seven CTAs per SM. (Thread-
//     private accumulators - one LDS.128 / STS.128 per row, no atomics - were built and measured first: 32 KB per
//     CTA, four CTAs per SM, 52 % issue utilisation and slower than lean2g: profiles/r02_lean2k_private_ncu.txt);
//   * sums are exact integers (value * 1000) as everywhere on the lean kernels; 5..7-byte decimals take an
//     out-of-line 64-bit route;
//   * "first row of the group" only costs something in the tile that numbers the group (the `fresh` bit).
// At most kL2KGroups groups per CTA: one more raises KERR_LEAN_GROUPS and the host reruns the scan on
// lean2g_kernel (32 groups), then on the packed global table. Rows outside the repertoire are handed over one
// by one (def_rows), tiles that are not clean as a whole (def_tiles): never approximated.
// The kernel is compiled per query only (cqg_jit: the shape as macros); without the run-time compiler the plan
// runs on lean2g_kernel.
#pragma once
#include "cqg_lean2g.cuh"

namespace cqg {

constexpr int kL2KGroups = 16;      // groups a CTA numbers
constexpr int kL2KDictCap = 256;    // words, two per bucket: 0 empty, 2 being written, (hash & ~0x7f) | gid << 2 | fresh << 1 | 1
constexpr int kL2KKeyRec = 48;      // raw key 16 | tag 4, pad 4, first okey 8 | canonical w0 8, w1 8
// per-warp accumulators, [word][group] u32: count | sum lo[3] | sum hi[3] | NULL fields[2]
constexpr int kL2KAccWords = 9;
constexpr int kL2KWarpAcc = kL2KAccWords * kL2KGroups * 4;
constexpr uint32_t kL2KLock = 2u, kL2KFresh = 2u;

template <class G>
struct Lean2KLayout {
    static constexpr int OFF_MSK = G::STAGES * G::BUF;
    static constexpr int OFF_CMP = OFF_MSK + G::MASKW * 8;          // kMaxLeanLeaf x 4 fd x {lo, width | lo, width for '-' fields}
    static constexpr int OFF_MBAR = OFF_CMP + kMaxLeanLeaf * 64;
    static constexpr int OFF_KMASK = (OFF_MBAR + G::STAGES * 8 + 15) / 16 * 16;  // [17][4] words: the first `len` bytes of 16
    static constexpr int OFF_SCALE = OFF_KMASK + 17 * 16;           // 10^(3 - fd) at byte offset fd16
    static constexpr int OFF_DICT = OFF_SCALE + 64;
    static constexpr int OFF_KEYS = OFF_DICT + kL2KDictCap * 4;
    static constexpr int OFF_MISC = OFF_KEYS + kL2KGroups * kL2KKeyRec;  // ngroups | nfresh[2] | pad | fresh slots [2][16] u16
    static constexpr int OFF_ACC = OFF_MISC + 16 + 64;                    // [warp][word][group] u32
    static constexpr int TOTAL = OFF_ACC + G::NWARPS * kL2KWarpAcc;
    static_assert(OFF_ACC % 16 == 0 && OFF_KEYS % 16 == 0 && OFF_DICT % 8 == 0, "alignment");
};

#ifdef CQG_JIT

__device__ __forceinline__ void reds32(uint32_t a, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t atoms32(uint32_t a, uint32_t v) {
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(v) : "memory");
    return old;
}
// a decimal of 5..7 bytes (value * 1000 may not fit 32 bits) or one with a sign into the (lo, hi) words of a warp's sum
__device__ __noinline__ void l2k_add_big(uint32_t lo_addr, uint32_t hi_addr, uint32_t mant, uint32_t scale, bool negative) {
    unsigned long long v = (unsigned long long)mant * scale;
    if (negative) v = 0ull - v;  // (two's complement: the (lo, hi) words add up modulo 2^64)
    const uint32_t vlo = (uint32_t)v, old = atoms32(lo_addr, vlo);
    const uint32_t up = (uint32_t)(v >> 32) + ((old + vlo) < old ? 1u : 0u);
    if (up) reds32(hi_addr, up);
}

// decimal field outside the 4-byte route: mant | fd16 << 24 | state << 30 (0: a decimal of 5..7 bytes, or one with a
// leading sign - fd16 then carries 8 for '-', cqg_lean2.cuh: CQG_L2_SIGNED -, 1 empty = NULL, 2 anything else: the row is
// handed over)
__device__ __noinline__ uint32_t l2k_decode_slow(uint32_t fa, uint32_t len) {
    if (len == 0u) return 1u << 30;
    if (len > 7u) return 2u << 30;
    if (len > 4u) {
        const uint32_t r7 = lean2_dec7(fa, len);
        if (r7 >> 31) return r7 & 0x3fffffffu;  // mant < 2^24, fd16 (0x00..0x30) << 24
    }
    const uint32_t rs = lean2_signed(fa, len);
    if (rs >> 31) return rs & 0x3fffffffu;      // mant, fd16 | 8 * negative
    return 2u << 30;
}

// the row that starts at `okey >> 16` belongs to a group numbered in this very tile: it may come before the row
// that numbered it
__device__ __noinline__ void l2k_first_row(uint8_t* keyrec, uint64_t okey) {
    atomicMin((unsigned long long*)(keyrec + 24), (unsigned long long)okey);
}

// Find or number the group of a raw key (the first probe of the row loop missed). Returns the slot word, or
// 0xffffffff when the row must be handed over (its key has no canonical form here: dates, signed or long numbers,
// text that starts or ends with a blank), or when there is no group number left (the scan is then rerun elsewhere).
template <class LL>
__device__ __noinline__ uint32_t l2k_find_slow(uint8_t* smem, uint32_t sbase, uint32_t y0, uint32_t y1, uint32_t y2, uint32_t y3, uint32_t h,
                                               uint32_t fa, uint32_t len, int it, unsigned* errflags) {
    unsigned int* dict = (unsigned int*)(smem + LL::OFF_DICT);
    unsigned int* misc = (unsigned int*)(smem + LL::OFF_MISC);
    uint32_t i = (h >> 25) * 2u;  // first word of the home bucket
    for (int probes = 0; probes < kL2KDictCap;) {
        uint32_t cur = *(volatile unsigned int*)(dict + i);
        if (cur == 0u) {
            uint32_t tag;
            uint64_t w0, w1;
            if (!l2g_key_part(fa, len, sbase + LL::OFF_KMASK, tag, w0, w1)) return 0xffffffffu;
            cur = atomicCAS(dict + i, 0u, kL2KLock);
            if (cur == 0u) {
                const unsigned int id = atomicAdd(misc, 1u);
                if (id >= (unsigned)kL2KGroups) {
                    atomicOr(errflags, KERR_LEAN_ABORT | KERR_LEAN_GROUPS);
                    atomicExch(dict + i, 0u);
                    return 0xffffffffu;
                }
                uint8_t* kr = smem + LL::OFF_KEYS + id * kL2KKeyRec;
                *(uint4*)kr = make_uint4(y0, y1, y2, y3);
                *(uint32_t*)(kr + 16) = tag;
                *(uint64_t*)(kr + 24) = ~0ull;
                *(uint64_t*)(kr + 32) = w0;
                *(uint64_t*)(kr + 40) = w1;
                const unsigned int k = atomicAdd(misc + 1 + (it & 1), 1u);
                ((uint16_t*)(misc + 4))[16 * (it & 1) + (k & 15u)] = (uint16_t)i;
                __threadfence_block();
                cur = (h & ~0x7fu) | (id << 2) | kL2KFresh | 1u;
                atomicExch(dict + i, cur);
                return cur;
            }
        }
        if (cur == kL2KLock) continue;  // being written: look again
        if (((cur ^ h) & ~0x7fu) == 0u) {
            const uint4 k4 = *(const uint4*)(smem + LL::OFF_KEYS + ((cur >> 2) & 31u) * kL2KKeyRec);
            if (k4.x == y0 && k4.y == y1 && k4.z == y2 && k4.w == y3) return cur;
        }
        i = (i + 1) & (kL2KDictCap - 1);
        probes++;
    }
    atomicOr(errflags, KERR_LEAN_ABORT | KERR_LEAN_GROUPS);
    return 0xffffffffu;
}

// compile-time shape: which wanted slots hold numbers (read by a leaf or an aggregate)
__device__ constexpr uint32_t l2k_numeric_slots() {
    uint32_t m = 0;
    for (int c = 0; c < 6; c++)
        if (c < CQG_JIT_NLEAF) m |= 1u << CQG_JIT_LEAFSLOT(c);
    for (int a = 0; a < 4; a++)
        if (a < CQG_JIT_NAGG) m |= 1u << CQG_JIT_ASLOT(a);
    return m;
}

template <class G, int MINB>
__global__ void __launch_bounds__(G::THREADS, MINB) lean2k_kernel(const __grid_constant__ DevPlan P) {
    extern __shared__ __align__(128) uint8_t smem[];
    static_assert(G::STAGES == 1 && G::TILE == G::THREADS * 128, "one stage, 128 bytes per thread");
    using LL = Lean2KLayout<G>;
    constexpr int nwant = CQG_JIT_NWANT, nprog = CQG_JIT_NPROG, nagg = CQG_JIT_NAGG, kslot = CQG_JIT_GSLOT(0);
    constexpr int gap0 = CQG_JIT_GAP0, gap1 = CQG_JIT_GAP1, gap2 = CQG_JIT_GAP2, gap3 = CQG_JIT_GAP3;
    constexpr uint32_t numslots = l2k_numeric_slots();
    static_assert(CQG_JIT_NGC == 1 && nagg <= 3 && nwant >= 1 && nwant <= 4 && kslot >= 0 && kslot < nwant, "lean2k shape");
    uint32_t sbase;
    asm volatile("mov.u32 %0, %1;" : "=r"(sbase) : "r"(smem_u32(smem)));
    const uint32_t s_buf = sbase + G::OFF_BUF, s_msk = sbase + LL::OFF_MSK, s_cmp = sbase + LL::OFF_CMP;
    uint64_t* mbar = (uint64_t*)(smem + LL::OFF_MBAR);
    unsigned int* misc = (unsigned int*)(smem + LL::OFF_MISC);
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t s_acc = sbase + LL::OFF_ACC + (uint32_t)(tid >> 5) * kL2KWarpAcc;  // this warp's accumulators

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int w = G::BUF / 32 + tid; w < G::MASKW; w += G::THREADS) {
        sts32(s_msk + 8 * w, 0xffffffffu);
        sts32(s_msk + 8 * w + 4, 0u);
    }
    if (tid < P.l_nleaf * 4) {
        uint32_t lo, width, clo, cwidth;
        lean2_interval(P.l_leaf[tid >> 2], tid & 3, lo, width, clo, cwidth);
        sts64(s_cmp + 16 * tid, lo, width);
        lean2_interval_neg(P.l_leaf[tid >> 2], tid & 3, clo, cwidth);  // fields with a leading '-'
        sts64(s_cmp + 16 * tid + 8, clo, cwidth);
    }
    for (int k = tid; k < (LL::TOTAL - LL::OFF_DICT) / 4; k += G::THREADS) ((uint32_t*)(smem + LL::OFF_DICT))[k] = 0u;
    if (tid < 17 * 4) {
        const int nb = (tid >> 2) - 4 * (tid & 3);  // bytes of word (tid & 3) inside a text of (tid >> 2) bytes
        ((uint32_t*)(smem + LL::OFF_KMASK))[tid] = nb >= 4 ? 0xffffffffu : (nb <= 0 ? 0u : (1u << (8 * nb)) - 1u);
    }
    if (tid < 4) ((uint32_t*)(smem + LL::OFF_SCALE))[4 * tid] = tid == 0 ? 1000u : tid == 1 ? 100u : tid == 2 ? 10u : 1u;
    __syncthreads();

    uint32_t rows = 0;
    const uint64_t size = P.size;
    const uint32_t patD = (uint32_t)P.delim * 0x01010101u;
    const uint32_t one = (uint32_t)P.simple >> 1;  // simple == 2 here: 1, but not to the compiler (IMAD adds)
    constexpr bool cr_too = CQG_JIT_CRLF != 0;  // CR is a line terminator too (DevPlan::crlf; part of the compiled shape here)

    const int my_tiles = (P.n_tiles > (int)blockIdx.x) ? (P.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    for (int it = 0; it < my_tiles; it++) {
        const long long tile = (long long)P.first_tile + blockIdx.x + (long long)it * gridDim.x;
        const long long g0 = tile * (long long)G::TILE - G::PRE;
        const bool at_edge = g0 < 0 || g0 + G::BUF > (long long)size;
        const bool edge = at_edge && !P.edge_in_kernel;  // handed over, not loaded (cqg_lean2.cuh: LeanEdge)
        LeanEdge es{0u, 0u, 0u};
        if (at_edge) es = lean_edge_span<G>(g0, size);  // (two tiles of a scan: not worth a dozen instructions on every tile)
        if (tid == 0) {
            if (!at_edge) {
                mbar_expect_tx(&mbar[0], G::BUF);
                tma_load_1d(smem + G::OFF_BUF, P.data + g0, G::BUF, &mbar[0]);
            } else if (!edge && es.load16) {
                mbar_expect_tx(&mbar[0], es.load16);
                tma_load_1d(smem + G::OFF_BUF + es.lo_b, P.data + g0 + (long long)es.lo_b, es.load16, &mbar[0]);
            } else {
                mbar_expect_tx(&mbar[0], 0);
            }
            const long long gn = g0 + (long long)gridDim.x * G::TILE;
            if (it + 1 < my_tiles && gn + G::BUF <= (long long)size)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(P.data + gn), "r"((uint32_t)G::BUF) : "memory");
        }
        // groups numbered in the previous tile are no longer fresh (their first row cannot lie in this tile)
        if (it > 0) {
            const unsigned nf = misc[1 + ((it - 1) & 1)];
            if ((unsigned)tid < (nf < 16u ? nf : 16u))
                atomicAnd((unsigned int*)(smem + LL::OFF_DICT) + ((const uint16_t*)(misc + 4))[16 * ((it - 1) & 1) + tid], ~kL2KFresh);
        }
        if (tid == 0) misc[1 + (it & 1)] = 0u;
        const unsigned abort_now = (*(volatile unsigned*)P.errflags) & KERR_LEAN_ABORT;
        mbar_wait(&mbar[0], (uint32_t)it & 1u);
        if (at_edge && !edge) {
            lean_edge_fill<G>(smem + G::OFF_BUF, P.data, g0, es, tid);
            __syncthreads();
        }

        // ---- phase 1: '\n' and delimiter masks, and "is the tile clean" (no other control byte, no quote) ----
        uint32_t spec = edge ? 0x80u : 0u;
        if (!edge) {
            const uint32_t ca0 = s_buf + 16u * tid;
            const uint32_t ma0 = s_msk + (((uint32_t)tid >> 1) << 3) + (((uint32_t)tid & 1u) << 1);
            auto chunk = [&](uint32_t ca, uint32_t ma) {
                const uint4 v = lds128(ca);
                uint32_t ra, rd;
                spec |= l2g_masks16(v.x, v.y, v.z, v.w, patD, one, ra, rd, cr_too);
                sts16(ma, ra);
                sts16(ma + 4u, rd);
            };
            constexpr int kFull = G::CHUNKS / G::THREADS;
#pragma unroll
            for (int k = 0; k < kFull; k++) chunk(ca0 + 16u * G::THREADS * k, ma0 + 4u * G::THREADS * k);
            if (tid < G::CHUNKS - kFull * G::THREADS) chunk(ca0 + 16u * G::THREADS * kFull, ma0 + 4u * G::THREADS * kFull);
            spec &= 0x80808080u;
        }
        const int special = __syncthreads_or((int)(spec != 0u) | (int)abort_now);
        if (special) {
            if (tid == 0 && !abort_now) {
                unsigned long long k = atomicAdd(P.def_tile_count, 1ull);
                P.def_tiles[k] = (int32_t)tile;
            }
            __syncthreads();
            continue;
        }

        // ---- phase 2: every thread walks the rows that START in its own 128 bytes ----
        uint32_t lo = (uint32_t)G::PRE + 128u * (uint32_t)tid, hi = lo + 128u;
        if (it == 0 || it + 1 == my_tiles) {  // only the scan's first and last tile can hold an end of the ownership range
            const long long olo_l = (long long)P.own_lo - g0, ohi_l = (long long)P.own_hi - g0;
            if (olo_l > (long long)G::PRE || ohi_l < (long long)(G::PRE + G::TILE)) {
                const uint32_t olo = (uint32_t)(olo_l < G::PRE ? G::PRE : (olo_l > G::PRE + G::TILE ? G::PRE + G::TILE : olo_l));
                const uint32_t ohi = (uint32_t)(ohi_l < G::PRE ? G::PRE : (ohi_l > G::PRE + G::TILE ? G::PRE + G::TILE : ohi_l));
                lo = lo > olo ? lo : olo;
                hi = hi < ohi ? hi : ohi;
            }
        }
        uint32_t handed = 0, myrows = 0;
        if (lo < hi) {
            // the first row start at or after lo: the byte behind the first '\n' at or after lo - 1
            uint32_t pos = lean2_next_term(s_msk, lo - 1u, (uint32_t)G::BUF) + 1u;
            if (pos < hi) {
                uint32_t ma = s_msk + ((pos >> 2) & ~7u);
                uint2 m0 = lds64(ma), m1 = lds64(ma + 8u);
                do {
                    const uint32_t tw = __funnelshift_r(m0.x, m1.x, pos);
                    const uint32_t dw = __funnelshift_r(m0.y, m1.y, pos);
                    uint32_t et, bad = 0u;
                    uint32_t off[4] = {0u, 0u, 0u, 0u}, len[4] = {0u, 0u, 0u, 0u};
                    if (tw != 0u && (tw & 1u) == 0u) {
                        // the row ends inside the 32-bit window: its delimiters, its '\n' and every bit above it are the
                        // stops (a field the row does not have comes out empty, or 32 bytes long: handed over)
                        const uint32_t below = tw ^ (tw - 1u);
                        et = bfind32(below);
                        uint32_t st = dw | tw | ~below, sp = 0u;
#define CQG_L2K_FIELD(K, GAP)                          \
    if (nwant > K) {                                   \
        if (GAP > 0) {                                 \
            _Pragma("unroll") for (int i = 1; i < GAP; i++) st &= st - 1u; \
            sp = l2_ffs32(st);                         \
            st &= st - 1u;                             \
        }                                              \
        off[K] = sp;                                   \
        len[K] = ctz32(st) - sp;                       \
    }
                        CQG_L2K_FIELD(0, gap0)
                        CQG_L2K_FIELD(1, gap1)
                        CQG_L2K_FIELD(2, gap2)
                        CQG_L2K_FIELD(3, gap3)
#undef CQG_L2K_FIELD
                    } else if (tw & 1u) {
                        // an empty line is not a row
                        pos++;
                        ma = s_msk + ((pos >> 2) & ~7u);
                        m0 = lds64(ma);
                        m1 = lds64(ma + 8u);
                        continue;
                    } else {
                        const uint32_t t2 = lds32(ma + 16u);
                        const uint32_t tw2 = __funnelshift_r(m1.x, t2, pos);
                        if (tw2 != 0u) {
                            const L2GWide wr = l2g_wide_fields(tw2, dw, __funnelshift_r(m1.y, lds32(ma + 20u), pos), nwant, gap0, gap1, gap2, gap3);
                            et = wr.et;
                            off[0] = (uint32_t)wr.fields & 0xffu;
                            len[0] = ((uint32_t)wr.fields >> 8) & 0xffu;
                            off[1] = ((uint32_t)wr.fields >> 16) & 0xffu;
                            len[1] = (uint32_t)wr.fields >> 24;
                            off[2] = (uint32_t)(wr.fields >> 32) & 0xffu;
                            len[2] = ((uint32_t)(wr.fields >> 32) >> 8) & 0xffu;
                            off[3] = (uint32_t)(wr.fields >> 32) >> 16 & 0xffu;
                            len[3] = (uint32_t)(wr.fields >> 32) >> 24;
                        } else {
                            // 64 bytes or more: hand the row over; its end is where the walk goes on
                            const uint32_t e = lean2_next_term(s_msk, pos + 64u, (uint32_t)G::BUF);
                            et = e - pos;  // (e == BUF: no row of this thread starts behind it)
                            bad = 1u;
                        }
                    }
                    const uint32_t rbase = s_buf + pos;
                    // the next row's mask words: asked for now, used after the decode
                    // (CR LF files: the LF right behind the CR is skipped here rather than by a trip through the loop as an
                    // empty line)
                    const uint32_t npos = pos + et + 1u + (cr_too ? (((tw >> 1) >> (et & 31u)) & 1u) : 0u);
                    const uint32_t nma = s_msk + ((npos >> 2) & ~7u);
                    const uint2 n0 = lds64(nma), n1 = lds64(nma + 8u);
                    myrows++;

                    // ---- decimals: every slot a leaf or an aggregate reads, once ----
                    uint32_t mant[4] = {0u, 0u, 0u, 0u}, fd16[4] = {0u, 0u, 0u, 0u}, state[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                    for (int s = 0; s < 4; s++) {
                        if (s < nwant && ((numslots >> s) & 1u)) {
                            state[s] = 2u;
                            if (len[s] - 1u < 4u && lean2_dec4(rbase + off[s] + len[s], len[s], mant[s], fd16[s])) state[s] = 0u;
                            if (state[s]) {
                                const uint32_t r = l2k_decode_slow(rbase + off[s], len[s]);
                                mant[s] = r & 0x00ffffffu;
                                fd16[s] = (r >> 24) & 0x38u;
                                state[s] = (r >> 30) == 0u ? 3u : (r >> 30);  // 3: a decimal of 5..7 bytes, or one with a sign
                            }
                        }
                    }
                    // ---- WHERE and the SUM / AVG operands (value * 1000, exact), written for the common row: every operand
                    // a decimal of <= 4 bytes (state 0). Anything else is put right in the cold block below. ----
                    bool pass = true;
                    if (nprog) {
                        uint32_t bs = 0;
#pragma unroll
                        for (int pc = 0; pc < nprog; pc++) {
                            const int c = CQG_JIT_PROG(pc);
                            if (c >= 0) {
                                const int sl = CQG_JIT_LEAFSLOT(c);
                                const uint2 iv = lds64(s_cmp + 64u * (uint32_t)c + fd16[sl]);
                                bs = (bs << 1) | (mant[sl] - iv.x <= iv.y ? 1u : 0u);
                            } else if (c == -1) {
                                bs = ((bs >> 1) & ~1u) | ((bs >> 1) & bs & 1u);
                            } else if (c == -2) {
                                bs = ((bs >> 1) & ~1u) | (((bs >> 1) | bs) & 1u);
                            } else {
                                bs ^= 1u;
                            }
                        }
                        pass = (bs & 1u) != 0u;
                    }
                    uint32_t add[3] = {0u, 0u, 0u}, nulls = 0u, big = 0u;
#pragma unroll
                    for (int a = 0; a < 3; a++) {
                        if (a < nagg) {
                            const int sl = CQG_JIT_ASLOT(a);
                            add[a] = mant[sl] * lds32(sbase + LL::OFF_SCALE + fd16[sl]);
                        }
                    }
                    if (state[0] | state[1] | state[2] | state[3]) {
#pragma unroll
                        for (int c = 0; c < 6; c++) {
                            if (c < CQG_JIT_NLEAF) {
                                const int sl = CQG_JIT_LEAFSLOT(c);
                                if ((state[sl] - 1u) < 2u) bad = 1u;  // NULL (1) or not a decimal (2) in a comparison: the general kernel's
                            }
                        }
#pragma unroll
                        for (int a = 0; a < 3; a++) {
                            if (a < nagg) {
                                const int sl = CQG_JIT_ASLOT(a);
                                if (state[sl] == 1u) {
                                    if (nagg == 3) bad = 1u;  // (no room for NULL counts beside three sums)
                                    if (a < 2) nulls |= 1u << (16 * a);
                                    add[a] = 0u;
                                } else if (state[sl] == 3u) {
                                    big |= 1u << a;
                                    add[a] = 0u;
                                } else if (state[sl] == 2u) {
                                    bad = 1u;
                                }
                            }
                        }
                    }
                    // ---- GROUP BY: raw key bytes -> group number of this CTA ----
                    uint32_t cur;
                    {
                        const uint32_t fa = rbase + off[kslot], kl = len[kslot];
                        const uint32_t a = fa & ~3u, sh = fa << 3;
                        const uint32_t x0 = lds32(a), x1 = lds32(a + 4), x2 = lds32(a + 8), x3 = lds32(a + 12), x4 = lds32(a + 16);
                        const uint4 m = lds128(sbase + LL::OFF_KMASK + 16u * kl);  // (kl > 16: some other 16 bytes of this CTA's shared memory, and the row is handed over)
                        const uint32_t y0 = __funnelshift_r(x0, x1, sh) & m.x, y1 = __funnelshift_r(x1, x2, sh) & m.y;
                        const uint32_t y2 = __funnelshift_r(x2, x3, sh) & m.z, y3 = __funnelshift_r(x3, x4, sh) & m.w;
                        if (kl > 16u) bad = 1u;
                        const uint32_t h = y0 * 0x9E3779B1u + y1 * 0x85EBCA77u + y2 * 0xC2B2AE3Du + y3 * 0x27D4EB2Fu;
                        const uint2 bk = lds64(sbase + LL::OFF_DICT + ((h >> 22) & ~7u));  // the home bucket: two words
                        cur = ((bk.x ^ h) & ~0x7fu) == 0u ? bk.x : bk.y;
                        const uint4 k4 = lds128(sbase + LL::OFF_KEYS + (cur & 0x7cu) * (kL2KKeyRec / 4));
                        const uint32_t diff = ((cur ^ h) & ~0x7fu) | (k4.x ^ y0) | (k4.y ^ y1) | (k4.z ^ y2) | (k4.w ^ y3) | (~cur & 1u);
                        if (diff != 0u && bad == 0u && pass) {
                            cur = l2k_find_slow<LL>(smem, sbase, y0, y1, y2, y3, h, fa, kl, it, P.errflags);
                            if (cur == 0xffffffffu) bad = 1u;
                        }
                    }
                    if (bad) {
                        unsigned long long k = atomicAdd(P.def_row_count, 1ull);
                        if (k < P.def_row_cap) P.def_rows[k] = (uint64_t)(g0 + (long long)pos);
                        handed++;
                    } else {
                        if (pass) {
                            const uint32_t aa = s_acc + (cur & 0x7cu);  // word [0][group]
                            reds32(aa, 1u);
                            if ((big | nulls) == 0u) {
#pragma unroll
                                for (int a = 0; a < 3; a++) {
                                    if (a < nagg) {
                                        const uint32_t old = atoms32(aa + (1 + a) * (kL2KGroups * 4), add[a]);
                                        if (old + add[a] < old) reds32(aa + (4 + a) * (kL2KGroups * 4), 1u);
                                    }
                                }
                            } else {
#pragma unroll
                                for (int a = 0; a < 3; a++) {
                                    if (a < nagg) {
                                        const uint32_t lo_a = aa + (1 + a) * (kL2KGroups * 4), hi_a = aa + (4 + a) * (kL2KGroups * 4);
                                        if ((big >> a) & 1u) {
                                            const int sl = CQG_JIT_ASLOT(a);
                                            l2k_add_big(lo_a, hi_a, mant[sl], lds32(sbase + LL::OFF_SCALE + (fd16[sl] & 0x30u)), (fd16[sl] & 8u) != 0u);
                                        } else if (a < 2 && ((nulls >> (16 * a)) & 1u)) {
                                            reds32(aa + (7 + a) * (kL2KGroups * 4), 1u);  // a NULL field: not a value of this aggregate
                                        } else {
                                            const uint32_t old = atoms32(lo_a, add[a]);
                                            if (old + add[a] < old) reds32(hi_a, 1u);
                                        }
                                    }
                                }
                            }
                            if (cur & kL2KFresh)
                                l2k_first_row(smem + LL::OFF_KEYS + ((cur >> 2) & 31u) * kL2KKeyRec, (P.global_base + (uint64_t)(g0 + (long long)pos)) << 16);
                        }
                    }
                    pos = npos;
                    ma = nma;
                    m0 = n0;
                    m1 = n1;
                } while (pos < hi);
            }
        }
        // too many rows outside this kernel's repertoire: let the general kernel do the whole scan.
        // The barrier also keeps the tile and its masks alive until every thread is done with them.
        rows += myrows - handed;  // (rows this kernel took itself)
        const int many = __syncthreads_or((int)(handed * 8u > myrows + 8u));
        if (many && tid == 0) atomicOr(P.errflags, KERR_LEAN_ABORT);
    }

    // ---- epilogue: the groups of this CTA into the global table ----
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) rows += __shfl_xor_sync(0xffffffffu, rows, d);
    if (lane == 0 && rows) atomicAdd(P.rows_scanned, (unsigned long long)rows);
    __syncthreads();
    unsigned err = 0;
    const unsigned ng = misc[0] < (unsigned)kL2KGroups ? misc[0] : (unsigned)kL2KGroups;
    if ((unsigned)tid < ng) {
        const uint8_t* kr = smem + LL::OFF_KEYS + tid * kL2KKeyRec;
        auto word = [&](int w, int k) { return *(const uint32_t*)(smem + LL::OFF_ACC + w * kL2KWarpAcc + (k * kL2KGroups + tid) * 4); };
        unsigned long long c = 0;
        for (int w = 0; w < G::NWARPS; w++) c += word(w, 0);
        if (c != 0ull) {
            const uint32_t tags = *(const uint32_t*)(kr + 16);
            uint64_t kw[2 * CQG_MAX_GROUP_COLS];
            for (int k = 0; k < 2 * CQG_MAX_GROUP_COLS; k++) kw[k] = 0;
            kw[0] = *(const uint64_t*)(kr + 32);
            kw[1] = *(const uint64_t*)(kr + 40);
            const uint64_t h = key_hash_final(key_hash_step(0x243F6A8885A308D3ull + 1ull, tags, kw[0], kw[1]));
            uint8_t* ge = global_entry_for(P, h, tags, kw, err);
            if (ge) {
                atomicAdd((unsigned long long*)(ge + kOffCount), c);
                amin64((uint64_t*)(ge + kOffFirst), *(const uint64_t*)(kr + 24));
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    if (a < nagg) {
                        unsigned long long t3 = 0, tn = 0;
                        for (int w = 0; w < G::NWARPS; w++) {
                            t3 += ((unsigned long long)word(w, 4 + a) << 32) + word(w, 1 + a);
                            if (a < 2) tn += word(w, 7 + a);
                        }
                        tn = c - tn;  // rows of the group minus those whose field was NULL
                        if (tn) {
                            atomicAdd((unsigned long long*)(ge + P.aggs[P.l_agg[a]].off + 16), tn);
                            atomicAdd((unsigned long long*)(ge + P.aggs[P.l_agg[a]].off + 24), t3);
                        }
                    }
                }
            }
        }
    }
    if (err) atomicOr(P.errflags, err);
}

#endif  // CQG_JIT

}  // namespace cqg
