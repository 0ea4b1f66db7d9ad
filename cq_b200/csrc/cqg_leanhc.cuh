// cqg_leanhc.cuh — lean GROUP BY for MANY groups (DevPlan::simple == 2, DevPlan::lean_global): BASELINE config 3,
// `GROUP BY name, surname, age, height` with SUM / MIN / MAX / AVG over ~2 M keys (create_groups
// evaluator_aggregates.c:108-176 + src/evaluator.c:112-212, evaluate_aggregate :263-326).
// lean2g_kernel's tile pipeline (cqg_lean2g.cuh: 1-D TMA tile, dot-product masks, exact '\n' class and clean-tile
// test, cursor row walk, right-aligned 4-byte decimal decode, interval leaves) with the per-CTA dictionary
// replaced by find-or-insert in the PACKED global table (PackedLayout, cqg_plan.cuh; packed_find, cqg_lean.cuh):
//   * one 32..256-byte line per group, read in 32-byte chunks (256-bit loads): what bounds this kernel is the NUMBER
//     of uncoalesced accesses a row makes (measured: ~80 G per second chip-wide, loads and atomics alike), so hash,
//     first okey, tags and key words come in one round trip of 1..3 loads, `first okey` and the MIN/MAX incumbents
//     are compared before an atomic is sent (both only ever move one way, so a stale read can cost a redundant
//     atomic, never a missed one), and an aggregate over a GROUP BY column keeps no state at all (it is a function
//     of the key and the count: config 3's SUM(age), MIN/MAX/AVG(height) - see PackedLayout::agg_key);
//   * a 2 x 32-bit multiplicative hash for the packed table (the general table's 64-bit hash is computed once per
//     GROUP by expand_packed_kernel, not once per row);
//   * numeric key parts reuse the decimal decode of the operands (a column that is key AND operand, like `height`
//     in config 3, is decoded once per row in a kernel compiled for the query);
//   * COUNT is the line's count; SUM / AVG add value*1000 (exact); MIN / MAX keep (ordered double image, okey) with a
//     128-bit CAS, ties to the earliest row - the value the reference keeps (evaluator_aggregates.c:316-321).
// Rows outside the repertoire (NULL operands, text in a number-only key slot, signed or long numbers, dates, rows
// of 64 bytes and more) are handed over one by one, tiles that are not clean as a whole; the general kernel builds
// general entries for them, and merge_general_into_dense_kernel folds those into the expanded lines.
#pragma once
#include "cqg_lean2g.cuh"

namespace cqg {

template <class G>
struct LeanHCLayout {
    static constexpr int OFF_MSK = G::STAGES * G::BUF;
    static constexpr int OFF_CMP = OFF_MSK + G::MASKW * 8;
    static constexpr int OFF_MBAR = OFF_CMP + kMaxLeanLeaf * 64;
    static constexpr int OFF_KMASK = (OFF_MBAR + G::STAGES * 8 + 15) / 16 * 16;  // [17][4] words: the first `len` bytes of 16
    static constexpr int TOTAL = OFF_KMASK + 17 * 16;
};

// unsigned decimal of 5..7 bytes, out of line: mant (< 10^7) | has-dot << 27 | fd << 28 | ok << 31
__device__ __noinline__ uint32_t hc_dec7(uint32_t fa, uint32_t len) {
    uint32_t mant, fd;
    bool hd;
    const bool ok = lean_decimal(fa, len, mant, fd, hd);
    return ok ? (0x80000000u | (fd << 28) | (hd ? 0x08000000u : 0u) | mant) : 0u;
}

// unsigned decimal of 1..7 bytes at shared address rb + o: value = mant / 10^(fd16 / 16); `dot`: the text has a '.',
// i.e. the reference types it DOUBLE (src/csv_reader.c:133-193). false: not one.
__device__ __forceinline__ bool hc_decimal(uint32_t rb, uint32_t o, uint32_t l, uint32_t& mant, uint32_t& fd16, bool& dot) {
    if (l - 1u < 4u) {
        const uint32_t fe = rb + o + l, a = fe & ~3u;
        const uint32_t w = __funnelshift_r(lds32(a - 4u), lds32(a), fe << 3);  // bytes [fe-4, fe)
        uint32_t t = w ^ 0x30303030u;
        t &= ~(0x00ffffffu >> (8u * l - 8u));
        dot = (~((((t ^ 0x1e1e1e1eu) & 0x7f7f7f7fu) + 0x7f7f7f7fu) | t) & 0x80808080u) != 0u;
        return lean2_dec4_word(w, l, mant, fd16);
    }
    if (l - 1u < 7u) {
        const uint32_t r = hc_dec7(rb + o, l);
        mant = r & 0x00ffffffu;
        fd16 = (r >> 24) & 0x30u;
        dot = (r & 0x08000000u) != 0u;
        return (r >> 31) != 0u;
    }
    return false;
}

// per-slot cache of the decode in a kernel compiled for one query (the slots are then compile-time constants)
#ifdef CQG_JIT
#define CQG_HC_DECODE_STATE uint32_t hcache_mant[4] = {0u, 0u, 0u, 0u}, hcache_fd[4] = {0u, 0u, 0u, 0u}, hcache_state = 0u;
#define CQG_HC_DECODE(SL, RB, O, L, DEC, MANT, FD16, DOT)                                    \
    {                                                                                       \
        const uint32_t st_ = (hcache_state >> (4 * (SL))) & 7u;                             \
        if (st_ != 0u) {                                                                    \
            DEC = (st_ & 1u) != 0u;                                                         \
            DOT = (st_ & 4u) != 0u;                                                         \
            MANT = hcache_mant[SL];                                                         \
            FD16 = hcache_fd[SL];                                                           \
        } else {                                                                            \
            DEC = hc_decimal(RB, O, L, MANT, FD16, DOT);                                    \
            hcache_mant[SL] = MANT;                                                         \
            hcache_fd[SL] = FD16;                                                           \
            hcache_state |= ((DEC ? 1u : 2u) | (DOT ? 4u : 0u)) << (4 * (SL));              \
        }                                                                                   \
    }
#else
#define CQG_HC_DECODE_STATE
#define CQG_HC_DECODE(SL, RB, O, L, DEC, MANT, FD16, DOT) DEC = hc_decimal(RB, O, L, MANT, FD16, DOT);
#endif

// numeric extreme on a packed { key, okey } pair: num_extreme (cqg_scan.cuh) with the incumbent read in one access
__device__ __forceinline__ void pk_extreme(uint8_t* st /*16B aligned*/, uint64_t key, uint64_t okey, bool is_min) {
    uint64_t ck, co;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(ck), "=l"(co) : "l"(st) : "memory");
    for (;;) {
        const bool better = is_min ? (key < ck || (key == ck && okey < co)) : (key > ck || (key == ck && okey < co));
        if (!better) return;
        if (cas128(st, ck, co, key, okey)) return;
    }
}

#define CQG_PK_COUNTOFF CQG_SPEC(PKCOUNT, P.pk.count_off)
#define CQG_PK_AGGOFF(a) CQG_SPEC_AT(PKAGGOFF, a, P.pk.agg_off[a])
#define CQG_PK_AGGKEY(a) CQG_SPEC_AT(PKAGGKEY, a, P.pk.agg_key[a])

template <class G, int MINB>
__global__ void __launch_bounds__(G::THREADS, MINB) leanhc_kernel(const __grid_constant__ DevPlan P) {
    extern __shared__ __align__(128) uint8_t smem[];
    static_assert(G::STAGES == 1 && G::TILE == G::THREADS * 128, "one stage, 128 bytes per thread");
    using LL = LeanHCLayout<G>;
    uint32_t sbase;
    asm volatile("mov.u32 %0, %1;" : "=r"(sbase) : "r"(smem_u32(smem)));
    const uint32_t s_buf = sbase + G::OFF_BUF, s_msk = sbase + LL::OFF_MSK, s_cmp = sbase + LL::OFF_CMP;
    uint64_t* mbar = (uint64_t*)(smem + LL::OFF_MBAR);
    const int tid = threadIdx.x, lane = tid & 31;

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int w = G::BUF / 32 + tid; w < G::MASKW; w += G::THREADS) {
        sts32(s_msk + 8 * w, 0xffffffffu);
        sts32(s_msk + 8 * w + 4, 0u);
    }
    if (tid < P.l_nleaf * 4 && P.l_leaf[tid >> 2].kind == 0) {
        uint32_t lo, width, clo, cwidth;
        lean2_interval(P.l_leaf[tid >> 2], tid & 3, lo, width, clo, cwidth);
        sts64(s_cmp + 16 * tid, lo, width);
    }
    if (tid < 17 * 4) {
        const int nb = (tid >> 2) - 4 * (tid & 3);  // bytes of word (tid & 3) inside a text of (tid >> 2) bytes
        ((uint32_t*)(smem + LL::OFF_KMASK))[tid] = nb >= 4 ? 0xffffffffu : (nb <= 0 ? 0u : (1u << (8 * nb)) - 1u);
    }
    __syncthreads();

    uint32_t rows = 0;
    const uint64_t size = P.size;
    const uint32_t patD = (uint32_t)P.delim * 0x01010101u;
    const uint32_t one = (uint32_t)P.simple >> 1;  // simple == 2 here: 1, but not to the compiler (IMAD adds)
    const bool cr_too = P.crlf != 0;
    const int nwant = CQG_SPEC(NWANT, P.nwantL);
    const int gap0 = CQG_SPEC(GAP0, P.gap[0]), gap1 = CQG_SPEC(GAP1, P.gap[1]), gap2 = CQG_SPEC(GAP2, P.gap[2]),
              gap3 = CQG_SPEC(GAP3, P.gap[3]);
    const int nprog = CQG_SPEC(NPROG, P.l_nprog);
    const int ngc = CQG_SPEC(NGC, P.ngc);
    const int nagg = CQG_SPEC(NAGG, P.l_nagg);
    uint32_t summask = 0;  // bits 0..3: aggregates with a state in the line; 4..7: those that are MIN/MAX; 8..11: MIN
    uint32_t numkeys = 0;  // key parts that must be numbers: an aggregate is derived from them (PackedLayout::agg_key)
    int aslot[4];
#pragma unroll
    for (int a = 0; a < 4; a++) {
        aslot[a] = 0;
        if (a < nagg) {
            const int func = CQG_SPEC_AT(AFUNC, a, P.aggs[P.l_agg[a]].func);
            const int kg = CQG_PK_AGGKEY(a);
            if (kg >= 0) {
                numkeys |= 1u << kg;
            } else {
                summask |= 1u << a;
                if (func == CQG_AGG_MIN || func == CQG_AGG_MAX) summask |= 16u << a;
                if (func == CQG_AGG_MIN) summask |= 256u << a;
                aslot[a] = CQG_SPEC_AT(ASLOT, a, P.aggs[P.l_agg[a]].slot);
            }
        }
    }

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
        const unsigned abort_now = (*(volatile unsigned*)P.errflags) & (KERR_LEAN_ABORT | KERR_TABLE_FULL);
        mbar_wait(&mbar[0], (uint32_t)it & 1u);
        if (at_edge && !edge) {
            lean_edge_fill<G>(smem + G::OFF_BUF, P.data, g0, es, tid);
            __syncthreads();
        }

        // ---- phase 1: '\n' and delimiter masks, and "is the tile clean" (no other byte below 0x23) ----
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
        const int special = __syncthreads_or((int)(spec != 0u) | (int)(abort_now != 0u));
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
        {
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
            uint32_t pos = lean2_next_term(s_msk, lo - 1u, (uint32_t)G::BUF) + 1u;
            while (pos < hi) {
                const uint32_t ma = s_msk + ((pos >> 2) & ~7u);
                const uint2 m0 = lds64(ma), m1 = lds64(ma + 8u);
                const uint32_t tw = __funnelshift_r(m0.x, m1.x, pos);
                const uint32_t dw = __funnelshift_r(m0.y, m1.y, pos);
                if (tw & 1u) {  // an empty line is not a row
                    pos++;
                    continue;
                }
                bool ok = true, pass = true;
                uint32_t et;
                uint32_t off0 = 0, off1 = 0, off2 = 0, off3 = 0, len0 = 0, len1 = 0, len2 = 0, len3 = 0;
                if (tw != 0u) {
                    const uint32_t below = tw ^ (tw - 1u);
                    et = bfind32(below);
                    Lean2Stops<uint32_t> S{(dw | tw) & below, 0u, false};
                    S.field(gap0, off0, len0);
                    if (nwant > 1) S.field(gap1, off1, len1);
                    if (nwant > 2) S.field(gap2, off2, len2);
                    if (nwant > 3) S.field(gap3, off3, len3);
                } else {
                    const uint32_t t2 = lds32(ma + 16u);
                    const uint32_t tw2 = __funnelshift_r(m1.x, t2, pos);
                    if (tw2 != 0u) {
                        const L2GWide wr = l2g_wide_fields(tw2, dw, __funnelshift_r(m1.y, lds32(ma + 20u), pos), nwant, gap0, gap1, gap2, gap3);
                        et = wr.et;
                        off0 = (uint32_t)wr.fields & 0xffu;
                        len0 = ((uint32_t)wr.fields >> 8) & 0xffu;
                        off1 = ((uint32_t)wr.fields >> 16) & 0xffu;
                        len1 = (uint32_t)wr.fields >> 24;
                        off2 = (uint32_t)(wr.fields >> 32) & 0xffu;
                        len2 = ((uint32_t)(wr.fields >> 32) >> 8) & 0xffu;
                        off3 = ((uint32_t)(wr.fields >> 32) >> 16) & 0xffu;
                        len3 = (uint32_t)(wr.fields >> 32) >> 24;
                    } else {
                        const uint32_t e = lean2_next_term(s_msk, pos + 64u, (uint32_t)G::BUF);
                        et = e - pos;
                        ok = false;  // 64 bytes or more
                    }
                }
                myrows++;
                const uint32_t rbase = s_buf + pos;
                unsigned long long add0 = 0, add1 = 0, add2 = 0, add3 = 0;
                CQG_HC_DECODE_STATE
#define CQG_HC_SLOT(SL, O, L)                                                    \
    const uint32_t O = SL == 0 ? off0 : SL == 1 ? off1 : SL == 2 ? off2 : off3; \
    const uint32_t L = SL == 0 ? len0 : SL == 1 ? len1 : SL == 2 ? len2 : len3;
                // ---- WHERE ----
                if (ok && nprog) {
                    uint32_t bs = 0;
                    CQG_SPEC_UNROLL
                    for (int pc = 0; pc < nprog; pc++) {
                        const int c = CQG_SPEC_AT(PROG, pc, P.l_prog[pc]);
                        if (c >= 0) {
                            const int sl = CQG_SPEC_AT(LEAFSLOT, c, P.l_leaf[c].slot), kind = CQG_SPEC_AT(LEAFKIND, c, P.l_leaf[c].kind);
                            CQG_HC_SLOT(sl, o, l)
                            bool bv = false;
                            if (kind == 0) {
                                uint32_t mant = 0, fd16 = 0;
                                bool dec = false, dot = false;
                                CQG_HC_DECODE(sl, rbase, o, l, dec, mant, fd16, dot)
                                if (dec) {
                                    const uint2 iv = lds64(s_cmp + 64u * (uint32_t)c + fd16);
                                    bv = mant - iv.x <= iv.y;
                                } else {
                                    ok = false;
                                }
                            } else {
                                uint32_t tag;
                                uint64_t w0, w1;
                                if (l == 0u) {
                                    bv = kind == 2;
                                } else if (l > 16u) {
                                    const uint32_t c0 = lds8(rbase + o);
                                    const bool ns = (c0 - 48u) <= 9u || c0 == '+' || c0 == '-' || c0 == '.';
                                    if (ns || c0 == ' ' || lds8(rbase + o + l - 1u) == ' ') ok = false;  // (trimmed by the reference)
                                    bv = kind == 2;
                                } else if (l2g_key_part(rbase + o, l, sbase + LL::OFF_KMASK, tag, w0, w1) && (tag == KT_STR || tag == KT_NULL)) {
                                    if (tag == KT_NULL) {
                                        w0 = 0x4c4c554eull;
                                        w1 = 0;
                                    }
                                    const bool eq = (uint32_t)P.l_leaf[c].slen == l && w0 == P.l_leaf[c].w0 && w1 == P.l_leaf[c].w1;
                                    bv = kind == 1 ? eq : !eq;
                                } else {
                                    ok = false;
                                }
                            }
                            bs = (bs << 1) | (bv ? 1u : 0u);
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
                // ---- operands: SUM / AVG value * 1000 (exact); MIN / MAX the ordered image of the double the reference parses ----
                if (ok && pass && summask) {
#define CQG_HC_AGG(A, ADD)                                                                               \
    if (summask & (1u << A)) {                                                                           \
        const int sl = aslot[A];                                                                         \
        CQG_HC_SLOT(sl, o, l)                                                                            \
        uint32_t mant = 0, fd16 = 0;                                                                     \
        bool dec = false, dot = false;                                                                   \
        CQG_HC_DECODE(sl, rbase, o, l, dec, mant, fd16, dot)                                             \
        if (dec) {                                                                                       \
            if (summask & (16u << A)) {                                                                  \
                const uint32_t fd = fd16 >> 4;                                                           \
                const double dv = mant < 10000u ? P.dec_table[fd * 10000u + mant] : (double)mant / kPow10[fd]; \
                ADD = num_key(dv);                                                                       \
            } else {                                                                                     \
                ADD = (unsigned long long)mant * (fd16 == 0u ? 1000u : fd16 == 16u ? 100u : fd16 == 32u ? 10u : 1u); \
            }                                                                                            \
        } else {                                                                                         \
            ok = false; /* NULL, text, date, signed or long number: the general kernel (the packed SUM  \
                           state counts no values, so NULL operands cannot stay here) */                 \
        }                                                                                                \
    }
                    CQG_HC_AGG(0, add0)
                    CQG_HC_AGG(1, add1)
                    CQG_HC_AGG(2, add2)
                    CQG_HC_AGG(3, add3)
#undef CQG_HC_AGG
                }
                // ---- GROUP BY: key parts as canon_part<true> builds them, then the packed line ----
                uint8_t* gentry = nullptr;
                uint64_t okey = 0;
                if (ok && pass) {
                    uint64_t kw[8];
                    uint32_t tags = 0;
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        kw[2 * g] = 0;
                        kw[2 * g + 1] = 0;
                        if (g < ngc && ok) {
                            const int sl = CQG_SPEC_AT(GSLOT, g, P.gslot[g]);
                            uint32_t tag = KT_NULL;
                            if (sl >= 0) {
                                CQG_HC_SLOT(sl, o, l)
                                const uint32_t c0 = l ? lds8(rbase + o) : 0u;
                                if ((c0 - 48u) <= 9u || c0 == '.') {
                                    uint32_t mant = 0, fd16 = 0;
                                    bool dec = false, dot = false;
                                    CQG_HC_DECODE(sl, rbase, o, l, dec, mant, fd16, dot)
                                    ok = dec;
                                    if (!dec) {
                                        tag = KT_STR;  // (handed over)
                                    } else if (!dot) {
                                        tag = KT_INT;
                                        kw[2 * g] = mant;
                                    } else {
                                        tag = KT_DBL_POS;  // |x| * 10^6 rounded = mant * 10^(6 - fd) exactly (mant < 10^7)
                                        kw[2 * g] = (uint64_t)mant * (fd16 == 0u ? 1000000u : fd16 == 16u ? 100000u : fd16 == 32u ? 10000u : 1000u);
                                    }
                                } else if (c0 == '+' || c0 == '-') {
                                    ok = false;
                                } else {
                                    ok = l2g_key_part(rbase + o, l, sbase + LL::OFF_KMASK, tag, kw[2 * g], kw[2 * g + 1]);
                                    // text in a narrow (number-only) key slot goes to the general kernel
                                    if (tag == KT_STR && !P.pk.key_wide[g]) ok = false;
                                }
                            }
                            tags |= tag << (4 * g);
                            // a SUM/AVG/MIN/MAX over this GROUP BY column is derived from the key: it must be a number
                            if (((numkeys >> g) & 1u) && tag != KT_INT && tag != KT_DBL_POS) ok = false;
                        }
                    }
                    if (ok && (P.hc_debug & 2)) {
                        rows += (uint32_t)(packed_hash(ngc, tags, kw) == 77ull);
                    } else if (ok) {
                        const uint64_t h = packed_hash(ngc, tags, kw);
                        uint64_t word0;
                        gentry = packed_find<true>(P, ngc, h, tags, kw, word0);
                        if (!gentry) {
                            atomicOr(P.errflags, KERR_TABLE_FULL);
                        } else {
                            okey = (P.global_base + (uint64_t)(g0 + (long long)pos)) << 16;
                            // word 0 = (first okey + bias) | tags: the low 16 bits are the same in every candidate
                            const uint64_t cand = (okey + kPkOkeyBias) | tags;
                            if (cand < word0) atomicMin((unsigned long long*)gentry, (unsigned long long)cand);
                        }
                    }
                }
#undef CQG_HC_SLOT
                if (!ok) {
                    unsigned long long k = atomicAdd(P.def_row_count, 1ull);
                    if (k < P.def_row_cap) P.def_rows[k] = (uint64_t)(g0 + (long long)pos);
                    handed++;
                } else {
                    rows++;
                    if (gentry && !(P.hc_debug & 1)) {
                        atomicAdd((unsigned long long*)(gentry + CQG_PK_COUNTOFF), 1ull);
#define CQG_HC_UPD(A, ADD)                                                                                 \
    if (summask & (1u << A)) {                                                                             \
        if (summask & (16u << A)) {                                                                        \
            pk_extreme(gentry + CQG_PK_AGGOFF(A), ADD, okey, (summask & (256u << A)) != 0u);               \
        } else {                                                                                           \
            atomicAdd((unsigned long long*)(gentry + CQG_PK_AGGOFF(A)), (unsigned long long)ADD);          \
        }                                                                                                  \
    }
                        CQG_HC_UPD(0, add0)
                        CQG_HC_UPD(1, add1)
                        CQG_HC_UPD(2, add2)
                        CQG_HC_UPD(3, add3)
#undef CQG_HC_UPD
                    }
                }
                pos += et + 1u + (((tw >> 1) >> (et & 31u)) & 1u);  // (+1: the LF of a CR LF pair, an empty line)
            }
        }
        // too many rows outside this kernel's repertoire: let the general kernel do the whole scan.
        // The barrier also keeps the tile and its masks alive until every thread is done with them.
        const int many = __syncthreads_or((int)(handed * 8u > myrows + 8u));
        if (many && tid == 0) atomicOr(P.errflags, KERR_LEAN_ABORT);
    }

#pragma unroll
    for (int d = 16; d > 0; d >>= 1) rows += __shfl_xor_sync(0xffffffffu, rows, d);
    if (lane == 0 && rows) atomicAdd(P.rows_scanned, (unsigned long long)rows);
}

}  // namespace cqg
