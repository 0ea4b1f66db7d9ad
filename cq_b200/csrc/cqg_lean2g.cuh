// cqg_lean2g.cuh — lean GROUP BY for few groups (DevPlan::simple == 2 without MIN/MAX, per-CTA dictionary
// mode): the operators of lean_kernel<GROUPED> (cqg_lean.cuh: create_groups evaluator_aggregates.c:108-176,
// evaluate_aggregate :263-326 for COUNT / SUM / AVG) on the cheaper parts of lean2_kernel (cqg_lean2.cuh):
//   * masks through dot products, the exact '\n' class and the clean-tile test of lean_kernel (a tile with any
//     other byte below 0x23 goes to the general kernel BEFORE anything is accumulated, so rows can be added to
//     the shared-memory accumulators as they are met);
//   * cursor row walk, stop bits with sentinels, right-aligned 4-byte decimal decode, interval leaves;
//   * a 32-bit hash for the per-CTA dictionary (the table hash of the general kernel is computed once per
//     group in the epilogue, not once per row), key compare on 16-byte shared loads;
//   * per-warp accumulators with native 32-bit shared atomics (carry into a high word), as lean_kernel.
// A CTA that meets more than kL2Groups groups raises KERR_LEAN_GROUPS: the host reruns the scan on
// lean_kernel's global-table mode. Rows outside the repertoire are handed over one by one (def_rows).
#pragma once
#include "cqg_lean2.cuh"

namespace cqg {

constexpr int kL2Groups = 32;     // groups a CTA numbers
// dictionary: 1024 one-word slots (so that two of <= 32 keys almost never share a home slot and a warp's 32
// look-ups finish in one probe together) + one key record per group
constexpr int kL2DictCap = 1024;  // slots: 0 empty, 2 being written, (hash & ~0x7f) | gid << 1 | 1 ready (gid 63: overflow)
constexpr int kL2KeyRec = 80;     // tags 4 | tile of insertion 4 | first okey 8 | 4 x key part 16
constexpr uint32_t kL2Lock = 2u;
// per-warp accumulators: count u32[G] | 4 x { lo[G] hi[G] nulls[G] } (values summed = count - nulls)
constexpr int kL2AggBlock = kL2Groups * 12;
constexpr int kL2WarpAcc = kL2Groups * 4 + 4 * kL2AggBlock;

template <class G>
struct Lean2GLayout {
    static constexpr int OFF_MSK = G::STAGES * G::BUF;
    static constexpr int OFF_CMP = OFF_MSK + G::MASKW * 8;
    static constexpr int OFF_MBAR = OFF_CMP + kMaxLeanLeaf * 64;
    static constexpr int OFF_DICT = (OFF_MBAR + G::STAGES * 8 + 15) / 16 * 16;
    static constexpr int OFF_KEYS = OFF_DICT + kL2DictCap * 4;
    static constexpr int OFF_NG = OFF_KEYS + kL2Groups * kL2KeyRec;
    static constexpr int OFF_KMASK = OFF_NG + 16;   // [17][4] words: the first `len` bytes of 16
    static constexpr int OFF_ACC = OFF_KMASK + 17 * 16;
    static constexpr int TOTAL = OFF_ACC + G::NWARPS * kL2WarpAcc;
};

__device__ __forceinline__ uint32_t l2g_hash_word(uint32_t h, uint32_t x) { return (h ^ x) * 0x9E3779B1u; }

// phase 1 on one 16-byte chunk: bit i of `n16` = byte i is '\n', of `d16` = byte i is the delimiter; the return value
// has a 0x80 bit set when the chunk holds any OTHER special byte (a control - CR, tab, NUL ... - or the quote '"'; the
// blank is an ordinary byte; kLeanSpecialXor in cqg_lean.cuh): the tile is then left to the general kernel.
// (b ^ 0x02) - 0x21 per byte borrows into bit 7 exactly where the byte is special and bit 7 clear ('\n' bytes are
// lifted out of the way first); a borrow that crosses into the next byte can only ADD a flag.
// `cr_too` (DevPlan::crlf: the head of the file holds a CR): '\r' is a line terminator like '\n' instead of a special byte,
// exactly as csv_load splits lines (src/csv_reader.c:404-427: both end a line, empty lines are skipped) - a CR LF pair is
// a terminator followed by an empty line, which the row walks skip. Four more instructions per word, only for such files.
__host__ __device__ __forceinline__ uint32_t l2g_masks16(uint32_t vx, uint32_t vy, uint32_t vz, uint32_t vw, uint32_t patD, uint32_t one,
                                                         uint32_t& n16, uint32_t& d16, bool cr_too = false) {
    uint32_t f0 = ~(l2_add((vx ^ 0x0a0a0a0au) & 0x7f7f7f7fu, one, 0x7f7f7f7fu) | vx) & 0x80808080u;
    uint32_t f1 = ~(l2_add((vy ^ 0x0a0a0a0au) & 0x7f7f7f7fu, one, 0x7f7f7f7fu) | vy) & 0x80808080u;
    uint32_t f2 = ~(l2_add((vz ^ 0x0a0a0a0au) & 0x7f7f7f7fu, one, 0x7f7f7f7fu) | vz) & 0x80808080u;
    uint32_t f3 = ~(l2_add((vw ^ 0x0a0a0a0au) & 0x7f7f7f7fu, one, 0x7f7f7f7fu) | vw) & 0x80808080u;
    if (cr_too) {
        f0 |= ~(l2_add((vx ^ 0x0d0d0d0du) & 0x7f7f7f7fu, one, 0x7f7f7f7fu) | vx) & 0x80808080u;
        f1 |= ~(l2_add((vy ^ 0x0d0d0d0du) & 0x7f7f7f7fu, one, 0x7f7f7f7fu) | vy) & 0x80808080u;
        f2 |= ~(l2_add((vz ^ 0x0d0d0d0du) & 0x7f7f7f7fu, one, 0x7f7f7f7fu) | vz) & 0x80808080u;
        f3 |= ~(l2_add((vw ^ 0x0d0d0d0du) & 0x7f7f7f7fu, one, 0x7f7f7f7fu) | vw) & 0x80808080u;
    }
    const uint32_t d0 = ~(l2_add((vx ^ patD) & 0x7f7f7f7fu, one, 0x7f7f7f7fu) | vx) & 0x80808080u;
    const uint32_t d1 = ~(l2_add((vy ^ patD) & 0x7f7f7f7fu, one, 0x7f7f7f7fu) | vy) & 0x80808080u;
    const uint32_t d2 = ~(l2_add((vz ^ patD) & 0x7f7f7f7fu, one, 0x7f7f7f7fu) | vz) & 0x80808080u;
    const uint32_t d3 = ~(l2_add((vw ^ patD) & 0x7f7f7f7fu, one, 0x7f7f7f7fu) | vw) & 0x80808080u;
    const uint32_t x0 = (vx | f0) ^ kLeanSpecialXor, x1 = (vy | f1) ^ kLeanSpecialXor, x2 = (vz | f2) ^ kLeanSpecialXor,
                   x3 = (vw | f3) ^ kLeanSpecialXor;
    const uint32_t spec = (l2_add(x0, one, kLeanSpecialSub) & ~x0) | (l2_add(x1, one, kLeanSpecialSub) & ~x1) |
                          (l2_add(x2, one, kLeanSpecialSub) & ~x2) | (l2_add(x3, one, kLeanSpecialSub) & ~x3);
    uint32_t ra = l2_dp4a(f2, 0x08040201u, 0u);
    ra = l2_dp4a(f3, 0x80402010u, ra) * 256u;
    ra = l2_dp4a(f0, 0x08040201u, ra);
    ra = l2_dp4a(f1, 0x80402010u, ra);
    uint32_t rd = l2_dp4a(d2, 0x08040201u, 0u);
    rd = l2_dp4a(d3, 0x80402010u, rd) * 256u;
    rd = l2_dp4a(d0, 0x08040201u, rd);
    rd = l2_dp4a(d1, 0x80402010u, rd);
    n16 = ra >> 7;
    d16 = rd >> 7;
    return spec;
}

// cold paths, kept out of the row loop's instruction footprint
struct L2GKey {
    uint64_t w0, w1;
    uint32_t tag, ok;
};
__device__ __noinline__ L2GKey l2g_key_part_number(uint32_t fa, uint32_t len) {
    L2GKey k;
    k.ok = lean_key_part(fa, len, k.tag, k.w0, k.w1) ? 1u : 0u;
    return k;
}
// fields of a row of 32..63 bytes on 64-bit masks: (off | len << 8) of wanted field k in byte pair k, et on top
struct L2GWide {
    uint64_t fields;
    uint32_t et;
};
__device__ __noinline__ L2GWide l2g_wide_fields(uint32_t tw2, uint32_t dlo, uint32_t dhi, int nwant, int gap0, int gap1, int gap2,
                                               int gap3) {
    const uint64_t tw64 = (uint64_t)tw2 << 32;
    const uint64_t dw64 = ((uint64_t)dhi << 32) | dlo;
    const uint64_t below = tw64 ^ (tw64 - 1ull);
    L2GWide r;
    r.et = 32u + bfind32((uint32_t)(below >> 32));
    Lean2Stops<uint64_t> S{(dw64 | tw64) & below, 0u, false};
    uint32_t o0 = 0, o1 = 0, o2 = 0, o3 = 0, l0 = 0, l1 = 0, l2 = 0, l3 = 0;
    S.field(gap0, o0, l0);
    if (nwant > 1) S.field(gap1, o1, l1);
    if (nwant > 2) S.field(gap2, o2, l2);
    if (nwant > 3) S.field(gap3, o3, l3);
    r.fields = (uint64_t)(o0 | (l0 << 8) | (o1 << 16) | (l1 << 24)) | ((uint64_t)(o2 | (l2 << 8) | (o3 << 16) | (l3 << 24)) << 32);
    return r;
}

// lean_key_part (cqg_lean.cuh) with the text case done on 32-bit words and a byte-mask table in shared memory
// (`s_kmask`); numbers as keys take lean_key_part itself. Same (tag, w0, w1) as canon_part<true> builds.
__device__ __forceinline__ bool l2g_key_part(uint32_t fa, uint32_t len, uint32_t s_kmask, uint32_t& tag, uint64_t& w0, uint64_t& w1) {
    w0 = 0;
    w1 = 0;
    if (len == 0u) {
        tag = KT_NULL;
        return true;
    }
    const uint32_t c0 = lds8(fa);
    if ((c0 - 48u) <= 9u || c0 == '+' || c0 == '-' || c0 == '.') {
        const L2GKey k = l2g_key_part_number(fa, len);
        tag = k.tag;
        w0 = k.w0;
        w1 = k.w1;
        return k.ok != 0u;
    }
    if (len > 16u) return false;
    if (c0 == ' ' || lds8(fa + len - 1u) == ' ') return false;  // trimmed by the reference: the general kernel's business
    const uint32_t a = fa & ~3u, sh = fa << 3;
    const uint32_t x0 = lds32(a), x1 = lds32(a + 4), x2 = lds32(a + 8), x3 = lds32(a + 12), x4 = lds32(a + 16);
    const uint4 m = lds128(s_kmask + 16u * len);
    const uint32_t y0 = __funnelshift_r(x0, x1, sh) & m.x, y1 = __funnelshift_r(x1, x2, sh) & m.y;
    const uint32_t y2 = __funnelshift_r(x2, x3, sh) & m.z, y3 = __funnelshift_r(x3, x4, sh) & m.w;
    if (len == 4u && y0 == 0x4c4c554eu) {  // the text NULL is the NULL group
        tag = KT_NULL;
        return true;
    }
    tag = KT_STR;
    w0 = ((uint64_t)y1 << 32) | y0;
    w1 = ((uint64_t)y3 << 32) | y2;
    return true;
}

template <class G, int MINB>
__global__ void __launch_bounds__(G::THREADS, MINB) lean2g_kernel(const __grid_constant__ DevPlan P) {
    extern __shared__ __align__(128) uint8_t smem[];
    static_assert(G::STAGES == 1 && G::TILE == G::THREADS * 128, "one stage, 128 bytes per thread");
    using LL = Lean2GLayout<G>;
    uint32_t sbase;
    asm volatile("mov.u32 %0, %1;" : "=r"(sbase) : "r"(smem_u32(smem)));
    const uint32_t s_buf = sbase + G::OFF_BUF, s_msk = sbase + LL::OFF_MSK, s_cmp = sbase + LL::OFF_CMP;
    uint64_t* mbar = (uint64_t*)(smem + LL::OFF_MBAR);
    uint8_t* dict = smem + LL::OFF_DICT;
    uint8_t* keys = smem + LL::OFF_KEYS;
    unsigned int* ngroups = (unsigned int*)(smem + LL::OFF_NG);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t* wacc = smem + LL::OFF_ACC + warp * kL2WarpAcc;

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
        lean2_interval_neg(P.l_leaf[tid >> 2], tid & 3, clo, cwidth);  // fields with a leading '-' (cqg_lean2.cuh: CQG_L2_SIGNED)
        sts64(s_cmp + 16 * tid + 8, clo, cwidth);
    }
    {
        const int words = (LL::TOTAL - LL::OFF_DICT) / 4;
        for (int k = tid; k < words; k += G::THREADS) ((uint32_t*)dict)[k] = 0u;
    }
    __syncthreads();
    if (tid < 17 * 4) {
        const int nb = (tid >> 2) - 4 * (tid & 3);  // bytes of word (tid & 3) inside a text of (tid >> 2) bytes
        ((uint32_t*)(smem + LL::OFF_KMASK))[tid] = nb >= 4 ? 0xffffffffu : (nb <= 0 ? 0u : (1u << (8 * nb)) - 1u);
    }
    for (int k = tid; k < kL2Groups; k += G::THREADS) *(uint64_t*)(keys + k * kL2KeyRec + 8) = ~0ull;  // first okey
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
    uint32_t summask = 0;
    int aslot[4];
#pragma unroll
    for (int a = 0; a < 4; a++) {
        aslot[a] = 0;
        if (a < nagg) {
            summask |= 1u << a;
            aslot[a] = CQG_SPEC_AT(ASLOT, a, P.aggs[P.l_agg[a]].slot);
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
        const unsigned abort_now = (*(volatile unsigned*)P.errflags) & KERR_LEAN_ABORT;
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
            // the first row start at or after lo: the byte behind the first '\n' at or after lo - 1 ...
            uint32_t pos = lean2_next_term(s_msk, lo - 1u, (uint32_t)G::BUF) + 1u;
            while (pos < hi) {
                const uint32_t ma = s_msk + ((pos >> 2) & ~7u);
                const uint2 m0 = lds64(ma), m1 = lds64(ma + 8u);
                const uint32_t tw = __funnelshift_r(m0.x, m1.x, pos);
                const uint32_t dw = __funnelshift_r(m0.y, m1.y, pos);
                if (tw & 1u) {  // ... that is not a '\n' itself (empty lines are not rows)
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
                        // 64 bytes or more: hand the row over; its end is where the walk goes on
                        const uint32_t e = lean2_next_term(s_msk, pos + 64u, (uint32_t)G::BUF);
                        et = e - pos;  // (e == BUF: no row of this thread starts behind it)
                        ok = false;
                    }
                }
                myrows++;
                const uint32_t rbase = s_buf + pos;
                unsigned long long add0 = 0, add1 = 0, add2 = 0, add3 = 0;
                uint32_t addmask = 0, gid = 0xffffffffu;
                CQG_L2_DECODE_STATE
#define CQG_L2G_SLOT(SL, O, L)                                                   \
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
                            CQG_L2G_SLOT(sl, o, l)
                            bool bv = false;
                            if (kind == 0) {
                                uint32_t mant = 0, fd16 = 0;
                                bool dec = false;
                                CQG_L2_DECODE(sl, rbase, o, l, dec, mant, fd16)
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
                // ---- SUM / AVG operands: value * 1000, exact ----
                if (ok && pass && summask) {
#define CQG_L2G_AGG(A, ADD)                                                                              \
    if (summask & (1u << A)) {                                                                           \
        const int sl = aslot[A];                                                                         \
        CQG_L2G_SLOT(sl, o, l)                                                                           \
        uint32_t mant = 0, fd16 = 0;                                                                     \
        bool dec = false;                                                                                \
        CQG_L2_DECODE(sl, rbase, o, l, dec, mant, fd16) \
        if (dec) {                                                                                       \
            ADD = lean2_times_1000(mant, fd16);                                                          \
            addmask |= 1u << A;                                                                          \
        } else if (l != 0u) {                                                                            \
            ok = false;                                                                                  \
        }                                                                                                \
    }
                    CQG_L2G_AGG(0, add0)
                    CQG_L2G_AGG(1, add1)
                    CQG_L2G_AGG(2, add2)
                    CQG_L2G_AGG(3, add3)
#undef CQG_L2G_AGG
                }
                // ---- GROUP BY: key -> group number of this CTA ----
                if (ok && pass) {
                    uint64_t kw[8];
                    uint32_t tags = 0, h32 = 0x85ebca6bu + (uint32_t)ngc;
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        kw[2 * g] = 0;
                        kw[2 * g + 1] = 0;
                        if (g < ngc && ok) {
                            const int sl = CQG_SPEC_AT(GSLOT, g, P.gslot[g]);
                            uint32_t tag = KT_NULL;
                            if (sl >= 0) {
                                CQG_L2G_SLOT(sl, o, l)
                                ok = l2g_key_part(rbase + o, l, sbase + LL::OFF_KMASK, tag, kw[2 * g], kw[2 * g + 1]);
                            }
                            tags |= tag << (4 * g);
                            h32 = l2g_hash_word(h32, (uint32_t)kw[2 * g] + tag);
                            h32 = l2g_hash_word(h32, (uint32_t)(kw[2 * g] >> 32));
                            h32 = l2g_hash_word(h32, (uint32_t)kw[2 * g + 1]);
                            h32 = l2g_hash_word(h32, (uint32_t)(kw[2 * g + 1] >> 32));
                        }
                    }
                    if (ok) {
                        h32 ^= h32 >> 15;
                        uint32_t i = (h32 >> 7) & (kL2DictCap - 1);
                        for (int probes = 0; probes < kL2DictCap;) {
                            const uint32_t sa = sbase + LL::OFF_DICT + 4u * i;
                            uint32_t cur = lds32(sa);
                            if (cur == 0u) {
                                cur = atomicCAS((unsigned int*)(dict + 4u * i), 0u, kL2Lock);
                                if (cur == 0u) {
                                    const unsigned int id = atomicAdd(ngroups, 1u);
                                    const uint32_t g6 = id < (unsigned)kL2Groups ? id : 63u;
                                    if (id < (unsigned)kL2Groups) {
                                        uint8_t* kr = keys + id * kL2KeyRec;
                                        *(uint32_t*)kr = tags;
                                        *(uint32_t*)(kr + 4) = (uint32_t)it;
#pragma unroll
                                        for (int g = 0; g < 4; g++) {
                                            *(uint64_t*)(kr + 16 + 16 * g) = kw[2 * g];
                                            *(uint64_t*)(kr + 24 + 16 * g) = kw[2 * g + 1];
                                        }
                                    }
                                    __threadfence_block();
                                    atomicExch((unsigned int*)(dict + 4u * i), (h32 & ~0x7fu) | (g6 << 1) | 1u);
                                    gid = g6;
                                    break;
                                }
                            }
                            if (cur == kL2Lock) continue;  // being written: look again
                            if (((cur ^ h32) & ~0x7fu) == 0u) {
                                const uint32_t g6 = (cur >> 1) & 63u;
                                if (g6 == 63u) {
                                    gid = 63u;
                                    break;
                                }
                                const uint32_t ka = sbase + LL::OFF_KEYS + g6 * kL2KeyRec;
                                bool same = lds32(ka) == tags;
#pragma unroll
                                for (int g = 0; g < 4; g++) {
                                    if (g < ngc) {
                                        const uint4 k4 = lds128(ka + 16u + 16u * g);
                                        same = same && k4.x == (uint32_t)kw[2 * g] && k4.y == (uint32_t)(kw[2 * g] >> 32) &&
                                               k4.z == (uint32_t)kw[2 * g + 1] && k4.w == (uint32_t)(kw[2 * g + 1] >> 32);
                                    }
                                }
                                if (same) {
                                    gid = g6;
                                    break;
                                }
                            }
                            i = (i + 1) & (kL2DictCap - 1);
                            probes++;
                        }
                        if (gid >= (uint32_t)kL2Groups) {
                            // more groups than a CTA numbers: rerun on the global table
                            atomicOr(P.errflags, KERR_LEAN_ABORT | KERR_LEAN_GROUPS);
                            gid = 0xffffffffu;
                        } else if (lds32(sbase + LL::OFF_KEYS + gid * kL2KeyRec + 4u) == (uint32_t)it) {
                            // first appearance: only rows of the tile in which the group was numbered can come before
                            // the row that numbered it (this CTA's tiles ascend)
                            const uint64_t okey = (P.global_base + (uint64_t)(g0 + (long long)pos)) << 16;
                            atomicMin((unsigned long long*)(keys + gid * kL2KeyRec + 8), (unsigned long long)okey);
                        }
                    }
                }
#undef CQG_L2G_SLOT
                if (!ok) {
                    unsigned long long k = atomicAdd(P.def_row_count, 1ull);
                    if (k < P.def_row_cap) P.def_rows[k] = (uint64_t)(g0 + (long long)pos);
                    handed++;
                } else {
                    rows++;
                    if (pass && gid != 0xffffffffu) {
                        atomicAdd((unsigned int*)(wacc + 4 * gid), 1u);
#define CQG_L2G_SUM(A, ADD)                                                                \
    if ((addmask >> A) & 1u) {                                                             \
        uint8_t* b = wacc + kL2Groups * 4 + A * kL2AggBlock;                               \
        const uint32_t vlo = (uint32_t)ADD, vhi = (uint32_t)(ADD >> 32);                   \
        const uint32_t old = atomicAdd((unsigned int*)(b + 4 * gid), vlo);                 \
        const uint32_t up = vhi + ((old + vlo) < old ? 1u : 0u);                           \
        if (up) atomicAdd((unsigned int*)(b + kL2Groups * 4 + 4 * gid), up);               \
    } else if ((summask >> A) & 1u) {                                                      \
        /* a NULL field: the rows that do NOT count towards this aggregate are the rare ones */ \
        atomicAdd((unsigned int*)(wacc + kL2Groups * 4 + A * kL2AggBlock + kL2Groups * 8 + 4 * gid), 1u); \
    }
                        CQG_L2G_SUM(0, add0)
                        CQG_L2G_SUM(1, add1)
                        CQG_L2G_SUM(2, add2)
                        CQG_L2G_SUM(3, add3)
#undef CQG_L2G_SUM
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

    // ---- epilogue: every dictionary entry: add up the warps' accumulators, fold into the global table ----
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) rows += __shfl_xor_sync(0xffffffffu, rows, d);
    if (lane == 0 && rows) atomicAdd(P.rows_scanned, (unsigned long long)rows);
    __syncthreads();
    unsigned err = 0;
    const unsigned ng = *ngroups < (unsigned)kL2Groups ? *ngroups : (unsigned)kL2Groups;
    for (int sidx = tid; sidx < (int)ng; sidx += G::THREADS) {
        const uint8_t* e = keys + sidx * kL2KeyRec;
        const uint32_t gid = (uint32_t)sidx;
        unsigned long long c = 0;
        for (int w = 0; w < G::NWARPS; w++) c += *(const uint32_t*)(smem + LL::OFF_ACC + w * kL2WarpAcc + 4 * gid);
        if (c == 0ull) continue;
        const uint32_t tags = *(const uint32_t*)e;
        uint64_t kw[2 * CQG_MAX_GROUP_COLS];
        uint64_t h = 0x243F6A8885A308D3ull + (uint64_t)ngc;
        for (int g = 0; g < 4; g++) {
            kw[2 * g] = *(const uint64_t*)(e + 16 + 16 * g);
            kw[2 * g + 1] = *(const uint64_t*)(e + 24 + 16 * g);
            if (g < ngc) h = key_hash_step(h, (tags >> (4 * g)) & 15u, kw[2 * g], kw[2 * g + 1]);
        }
        h = key_hash_final(h);
        uint8_t* ge = global_entry_for(P, h, tags, kw, err);
        if (!ge) continue;
        atomicAdd((unsigned long long*)(ge + kOffCount), c);
        amin64((uint64_t*)(ge + kOffFirst), *(const uint64_t*)(e + 8));
        for (int a = 0; a < 4; a++) {
            if (a < nagg) {
                unsigned long long t3 = 0, tn = 0;
                for (int w = 0; w < G::NWARPS; w++) {
                    const uint8_t* b = smem + LL::OFF_ACC + w * kL2WarpAcc + kL2Groups * 4 + a * kL2AggBlock;
                    t3 += ((unsigned long long)*(const uint32_t*)(b + kL2Groups * 4 + 4 * gid) << 32) + *(const uint32_t*)(b + 4 * gid);
                    tn += *(const uint32_t*)(b + kL2Groups * 8 + 4 * gid);
                }
                tn = c - tn;  // rows of the group minus those whose field was NULL
                if (tn) {
                    atomicAdd((unsigned long long*)(ge + P.aggs[P.l_agg[a]].off + 16), tn);
                    atomicAdd((unsigned long long*)(ge + P.aggs[P.l_agg[a]].off + 24), t3);
                }
            }
        }
    }
    if (err) atomicOr(P.errflags, err);
}

}  // namespace cqg
