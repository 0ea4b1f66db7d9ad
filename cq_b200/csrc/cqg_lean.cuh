// cqg_lean.cuh — the lean scan kernel: the same fused pipeline as cqg_scan.cuh (TMA tile ->
// SWAR masks -> row list -> one row per thread) stripped to what DevPlan::simple plans need:
// no GROUP BY, COUNT/SUM/AVG, WHERE empty or `column <op> decimal literal`, fields that are
// short decimals. Everything else is HANDED OVER, never approximated:
//   * a tile holding any byte below 0x23 other than '\n' (CR, quote, blank, NUL) or touching a
//     file edge goes on def_tiles;
//   * a row of 64 bytes or more, or with a wanted field that is not a decimal of <= 7 bytes,
//     goes on def_rows;
// and the general kernel processes those afterwards into the same group entry. When more than
// 1/8 of a tile's rows are handed over the kernel raises KERR_LEAN_ABORT and the host reruns the
// whole scan on the general kernel.
// All shared-memory traffic uses 32-bit shared addresses (ld.shared / st.shared).
#pragma once
#include "cqg_scan.cuh"

// ---- per-query specialisation (cqg_jit in cqg_api.cu) ----
// Compiled ahead of time the lean kernels read the plan's SHAPE (wanted columns and the delimiters between
// them, the leaf program, key and aggregate slots) from the kernel parameter. Compiled at run time for one
// query (NVRTC, -DCQG_JIT) the same source gets the shape as macros: loops unroll, slot selects and leaf-kind
// branches fold away. Values (literals, intervals, offsets) stay in the parameter either way.
#ifdef CQG_JIT
#define CQG_SPEC(NAME, RUNTIME) (CQG_JIT_##NAME)
#define CQG_SPEC_AT(NAME, I, RUNTIME) (CQG_JIT_##NAME(I))
#define CQG_SPEC_UNROLL _Pragma("unroll")
#else
#define CQG_SPEC(NAME, RUNTIME) (RUNTIME)
#define CQG_SPEC_AT(NAME, I, RUNTIME) (RUNTIME)
#define CQG_SPEC_UNROLL _Pragma("unroll 1")
#endif

namespace cqg {

__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// unsigned decimal of 1..7 bytes at shared address `fa`: value = mant / 10^fd. false: not one.
// `hasdot`: the text had a '.', i.e. the reference types it DOUBLE, not INTEGER.
__device__ __forceinline__ bool lean_decimal(uint32_t fa, uint32_t len, uint32_t& mant, uint32_t& fd, bool& hasdot) {
    const uint32_t a = fa & ~3u, sh = (fa & 3u) * 8u;
    const uint32_t w0 = lds32(a), w1 = lds32(a + 4);
    if (len <= 4u) {
        uint32_t w = __funnelshift_r(w0, w1, sh);
        w = len == 4u ? w : ((w << (8u * (4u - len))) | (0x30303030u >> (8u * len)));
        uint32_t dotf = ~((((w ^ 0x2e2e2e2eu) & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) & 0x80808080u;
        uint32_t t = w ^ 0x30303030u;
        fd = 0;
        hasdot = dotf != 0u;
        uint32_t ndig = len;
        if (dotf) {
            if (dotf & (dotf - 1u)) return false;
            const uint32_t j = (31u - __clz(dotf)) >> 3;
            fd = 3u - j;
            const uint32_t high = j == 3u ? 0u : (t & ~((1u << (8u * (j + 1u))) - 1u));
            t = ((t & ((1u << (8u * j)) - 1u)) << 8) | high;
            ndig = len - 1u;
        }
        if (ndig == 0u) return false;
        if (((t + 0x76767676u) | t) & 0x80808080u) return false;
        t = t * 10u + (t >> 8);
        t &= 0x00ff00ffu;
        mant = (t * 100u + (t >> 16)) & 0xffffu;
        return true;
    }
    // 5..7 bytes: digit by digit out of three registers
    const uint32_t w2 = lds32(a + 8);
    uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
    uint32_t m = 0, f = 0;
    bool dot = false, ok = true, digit = false;
#pragma unroll
    for (uint32_t k = 0; k < 7u; k++) {
        if (k < len) {
            const uint32_t c = (k < 4u ? (lo >> (8u * k)) : (hi >> (8u * (k - 4u)))) & 0xffu;
            const uint32_t d = c - 48u;
            if (d <= 9u) {
                m = m * 10u + d;
                digit = true;
                f += dot ? 1u : 0u;
            } else {
                ok = ok && c == '.' && !dot;
                dot = true;
            }
        }
    }
    mant = m;
    fd = f;
    hasdot = dot;
    return ok && digit && f <= 3u;  // more than 3 fraction digits: leave to the general kernel
}

// ---- lean GROUP BY: key parts of a clean tile ----
// One GROUP BY key part exactly as canon_part<true> builds it, for the two shapes this kernel
// covers: text of <= 16 bytes that cannot start a number, and unsigned decimals of <= 7 bytes.
// false: hand the row over (dates, signed or long numbers, long strings, "1-2" and the like).
__device__ __forceinline__ bool lean_key_part(uint32_t fa, uint32_t len, uint32_t& tag, uint64_t& w0, uint64_t& w1) {
    w0 = 0;
    w1 = 0;
    if (len == 0u) {
        tag = KT_NULL;
        return true;
    }
    const uint32_t c0 = lds8(fa);
    const bool numeric_start = (c0 - 48u) <= 9u || c0 == '+' || c0 == '-' || c0 == '.';
    if (!numeric_start) {
        if (len > 16u) return false;
        // a blank is an ordinary byte of a clean tile, but the reference trims the value (src/csv_reader.c:195-240):
        // a field that starts or ends with one is decoded by the general kernel
        if (c0 == ' ' || lds8(fa + len - 1u) == ' ') return false;
        const uint32_t a = fa & ~3u, sh = (fa & 3u) * 8u;
        const uint32_t x0 = lds32(a), x1 = lds32(a + 4), x2 = lds32(a + 8), x3 = lds32(a + 12), x4 = lds32(a + 16);
        uint64_t lo = ((uint64_t)__funnelshift_r(x1, x2, sh) << 32) | __funnelshift_r(x0, x1, sh);
        uint64_t hi = ((uint64_t)__funnelshift_r(x3, x4, sh) << 32) | __funnelshift_r(x2, x3, sh);
        if (len < 8u) {
            lo &= (1ull << (8u * len)) - 1ull;
            hi = 0;
        } else if (len < 16u) {
            hi &= (1ull << (8u * (len - 8u))) - 1ull;  // len == 8: mask 0
        }
        if (len == 4u && (uint32_t)lo == 0x4c4c554eu) {  // the text NULL is the NULL group
            tag = KT_NULL;
            return true;
        }
        tag = KT_STR;
        w0 = lo;
        w1 = hi;
        return true;
    }
    if (len > 7u || c0 == '+' || c0 == '-') return false;
    uint32_t mant, fd;
    bool hasdot;
    if (!lean_decimal(fa, len, mant, fd, hasdot)) return false;
    if (!hasdot) {
        tag = KT_INT;
        w0 = mant;
    } else {
        tag = KT_DBL_POS;  // |x|*10^6 rounded = mant * 10^(6-fd) exactly (mant < 10^7)
        w0 = (uint64_t)mant * (fd == 0u ? 1000000u : fd == 1u ? 100000u : fd == 2u ? 10000u : 1000u);
    }
    return true;
}

// ---- the PACKED global table of the lean GROUP BY in global mode (PackedLayout in cqg_plan.cuh) ----
// Every access to a line that is not shared with other lines costs a round trip to L2 and, measured, ~1/80 G of a
// second chip-wide whatever its width: the line is therefore read in 32-byte chunks (one 256-bit load each,
// LDG.E.256) and a row sends as few atomics as it can.
__device__ __forceinline__ uint64_t ld_acquire_u64(const void* p) {
    uint64_t v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
struct PkChunk {
    uint64_t a, b, c, d;
};
__device__ __forceinline__ PkChunk pk_load(const void* p) {
    PkChunk v;
#if defined(__CUDACC_VER_MAJOR__) && (__CUDACC_VER_MAJOR__ * 100 + __CUDACC_VER_MINOR__ < 1209)
    // a run-time compiler older than CUDA 12.9 (PTX 8.8) has no 256-bit vector load: two 128-bit halves of the same sector
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(v.a), "=l"(v.b) : "l"(p) : "memory");
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2+16];" : "=l"(v.c), "=l"(v.d) : "l"(p) : "memory");
#else
    asm volatile("ld.relaxed.gpu.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v.a), "=l"(v.b), "=l"(v.c), "=l"(v.d) : "l"(p) : "memory");
#endif
    return v;
}
// word j (0..11) of a line read as three chunks; folds to one register when j is known at compile time
__device__ __forceinline__ uint64_t pk_word(const PkChunk& c0, const PkChunk& c1, const PkChunk& c2, int j) {
    switch (j) {
        case 0: return c0.a;
        case 1: return c0.b;
        case 2: return c0.c;
        case 3: return c0.d;
        case 4: return c1.a;
        case 5: return c1.b;
        case 6: return c1.c;
        case 7: return c1.d;
        case 8: return c2.a;
        case 9: return c2.b;
        case 10: return c2.c;
        default: return c2.d;
    }
}

// layout values: compile-time constants in a kernel compiled for one query (cqg_jit), plan reads otherwise
#define CQG_PK_IDWORDS CQG_SPEC(PKIDW, P.pk.id_words)
#define CQG_PK_KEYWORD(g) CQG_SPEC_AT(PKKEYWORD, g, P.pk.key_word[g])
#define CQG_PK_KEYWIDE(g) CQG_SPEC_AT(PKKEYWIDE, g, P.pk.key_wide[g])
#define CQG_PK_BYTES CQG_SPEC(PKBYTES, P.pk.entry_bytes)

// hash of a group key for the packed table only (two 32-bit multiplicative lanes, one avalanche each): it places
// lines and pre-filters probes, identity is always the key compare. Never 0, bit 63 (the lock bit) clear.
__device__ __forceinline__ uint32_t packed_mix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint64_t packed_hash(int ngc, uint32_t tags, const uint64_t* kw) {
    uint32_t a = 0x85ebca6bu + tags, b = 0xc2b2ae35u ^ (uint32_t)ngc;
#pragma unroll
    for (int g = 0; g < 4; g++) {
        if (g < ngc) {
            a = (a ^ (uint32_t)kw[2 * g]) * 0x9E3779B1u;
            b = (b ^ (uint32_t)(kw[2 * g] >> 32)) * 0x85EBCA77u;
            a = (a ^ (uint32_t)kw[2 * g + 1]) * 0xC2B2AE3Du;
            b = (b ^ (uint32_t)(kw[2 * g + 1] >> 32)) * 0x27D4EB2Fu;
        }
    }
    const uint32_t x = packed_mix32(a ^ __funnelshift_l(b, b, 16));
    const uint32_t y = packed_mix32(b + a * 0x9E3779B1u);
    return ((((uint64_t)y << 32) | x) & 0x7fffffffffffffffull) | 1ull;
}

// word 0 of a line: 0 = empty, kPkBusy = being written, else (first okey + kPkOkeyBias) | tags (never below 2^16)
constexpr uint64_t kPkBusy = 1ull;
constexpr uint64_t kPkOkeyBias = 1ull << 16;  // okeys are (row offset << 16): + one offset unit keeps word 0 away from 0 / busy
// Key words are stored XOR this constant, so that the image of a key word is never the 0 an unwritten word reads as
// (except for the one key word equal to the constant itself, which takes the ordered path below).
constexpr uint64_t kPkKeyMask = 0xA5A5A5A55A5A5A5Bull;

// The chunks a lookup needs, then tags and key words against (tags, kw), compared without short-circuit.
// The loads are not ordered among themselves: a key word may have been read before the inserting thread wrote it,
// and then reads 0. `unsure`: the outcome rests on such a word - a mismatch against a 0, or (for the one key word
// whose image is 0) a match with one - and must be confirmed with ordered loads.
__device__ __forceinline__ bool packed_same_key(const DevPlan& P, int ngc, const uint8_t* e, uint32_t tags, const uint64_t* kw,
                                                uint64_t& word0, bool& unsure) {
    const int idw = CQG_PK_IDWORDS;
    const PkChunk c0 = pk_load(e);
    PkChunk c1 = {0, 0, 0, 0}, c2 = {0, 0, 0, 0};
    if (idw > 4) c1 = pk_load(e + 32);
    if (idw > 8) c2 = pk_load(e + 64);
    word0 = c0.a;
    // all key words at once: XOR and OR (two LOP3 per word), one test. Which words read as 0 only matters when the
    // outcome could rest on one: a mismatch (any of them unwritten yet?) or, for a match, a key image that is 0 itself
    uint64_t diff = 0;
    bool zero_image = false;
#pragma unroll
    for (int g = 0; g < 4; g++) {
        if (g < ngc) {
            const int j = CQG_PK_KEYWORD(g);
            const uint64_t i0 = kw[2 * g] ^ kPkKeyMask;
            diff |= pk_word(c0, c1, c2, j) ^ i0;
            zero_image |= i0 == 0ull;
            if (CQG_PK_KEYWIDE(g)) {
                const uint64_t i1 = kw[2 * g + 1] ^ kPkKeyMask;
                diff |= pk_word(c0, c1, c2, j + 1) ^ i1;
                zero_image |= i1 == 0ull;
            }
        }
    }
    const bool same = (uint32_t)(c0.a & 0xffffull) == tags && diff == 0ull;
    bool zero_miss = false;
    if (!same) {
#pragma unroll
        for (int g = 0; g < 4; g++) {
            if (g < ngc) {
                const int j = CQG_PK_KEYWORD(g);
                zero_miss |= (pk_word(c0, c1, c2, j) == 0ull) & ((kw[2 * g] ^ kPkKeyMask) != 0ull);
                if (CQG_PK_KEYWIDE(g)) zero_miss |= (pk_word(c0, c1, c2, j + 1) == 0ull) & ((kw[2 * g + 1] ^ kPkKeyMask) != 0ull);
            }
        }
    }
    unsure = same ? zero_image : zero_miss;
    return same;
}

// find-or-insert. The steady state is a lookup: the chunks holding word 0 (first okey | tags) and the key words are
// requested together (one memory round trip). A line is claimed by CAS 0 -> busy on word 0, its key words are
// written (each line once per query, over the 0 the table is initialised with), and it is published by storing
// (no first row yet) | tags over `busy` after a fence. A published word 0 with all key images equal is this key:
// every image is nonzero, so every word compared had been written. Anything resting on a word that read 0 is
// confirmed after an acquire load of word 0.
// INSERT = false: lookup only (merge of the general kernel's entries into the expanded ones).
// `word0`: as loaded (may be stale, i.e. too large: it only decides whether the atomicMin can be skipped).
template <bool INSERT>
__device__ __forceinline__ uint8_t* packed_find(const DevPlan& P, int ngc, uint64_t h, uint32_t tags, const uint64_t* kw, uint64_t& word0) {
    const uint64_t mask = P.pcap - 1;
    const uint32_t eb = (uint32_t)CQG_PK_BYTES;
    uint64_t i = (h >> 1) & mask;
    for (uint64_t probes = 0; probes < P.pcap;) {
        uint8_t* e = P.ptab + i * (uint64_t)eb;
        bool unsure;
        bool same = packed_same_key(P, ngc, e, tags, kw, word0, unsure);
        bool recheck = unsure;
        if (word0 == 0ull) {
            if (!INSERT) return nullptr;
            if (*(volatile unsigned long long*)P.pcount >= P.pcap / 2) return nullptr;
            word0 = atomicCAS((unsigned long long*)e, 0ull, (unsigned long long)kPkBusy);
            if (word0 == 0ull) {
                atomicAdd(P.pcount, 1ull);
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    if (g < ngc) {
                        const int j = CQG_PK_KEYWORD(g);
                        *(uint64_t*)(e + 8 * j) = kw[2 * g] ^ kPkKeyMask;
                        if (CQG_PK_KEYWIDE(g)) *(uint64_t*)(e + 8 * j + 8) = kw[2 * g + 1] ^ kPkKeyMask;
                    }
                }
                __threadfence();
                word0 = 0xffffffffffff0000ull | tags;
                atomicExch((unsigned long long*)e, (unsigned long long)word0);
                return e;
            }
            recheck = true;  // somebody else claimed the line meanwhile: what was loaded above is older than that
        }
        if (word0 == kPkBusy) {
            while (ld_acquire_u64(e) == kPkBusy) {
            }
            recheck = true;
        }
        if (recheck) {
            (void)ld_acquire_u64(e);  // word 0 is published: loads behind this one see the key words
            same = packed_same_key(P, ngc, e, tags, kw, word0, unsure);
        }
        if (same) return e;
        i = (i + 1) & mask;
        probes++;
    }
    return nullptr;
}

// per-CTA dictionary of the lean GROUP BY: key -> small group number
constexpr int kLeanGroups = 64;     // groups a CTA can number; one more aborts to the general kernel
constexpr int kLeanDictCap = 128;   // slots
constexpr int kLeanDictEntry = 96;  // hash 8 | gid 4 | tags 4 | first okey 8 | pad 8 | 4 x key part 16
// per-warp accumulators (only that warp writes them): count u32 [G] | 4 aggregate blocks
// aggregate block: SUM/AVG lo[G] hi[G] n[G] (u32); with MIN/MAX in the plan G x { fn u64, pad, key u64, okey u64 }
__host__ __device__ constexpr int lean_agg_block(bool minmax) { return kLeanGroups * (minmax ? 32 : 12); }
__host__ __device__ constexpr int lean_warp_acc(bool minmax) { return kLeanGroups * 4 + 4 * lean_agg_block(minmax); }

// The bytes a lean tile may not hold (other than '\n'): controls (< 0x20: CR, tab, NUL ...) and the quote '"' (0x22).
// The blank (0x20) and '!' (0x21) are ordinary bytes: a file with `New York` in it stays on the lean kernels. One
// test covers both ranges: b ^ 0x02 maps 0x22 -> 0x20 and 0x20, 0x21 -> 0x22, 0x23 and keeps every control below
// 0x20, so "special" is (b ^ 0x02) < 0x21 (and bit 7 clear). Plans whose quote character is anything else than '"'
// or a control take the general kernel (DevPlan::exact_only).
constexpr uint32_t kLeanSpecialXor = 0x02020202u;
constexpr uint32_t kLeanSpecialSub = 0xdededEdfu;  // -0x21212121

// a + c issued as IMAD (a * 1 + c) so that it runs on the FMA pipe: LOP3/IADD3/SHF all share the ALU pipe,
// which takes one warp instruction every two cycles; phase 1 is otherwise all-ALU (B300_MICROARCH: pipe rates)
__device__ __forceinline__ uint32_t add_fma(uint32_t a, uint32_t one, uint32_t c) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(c));
    return d;
}

// four flag words -> 16-bit mask in byte order: the multiplies leave the nibbles at bits 0..3 / 4..7 of the
// high word with garbage only above bit 7, so one select merges two of them and one byte-permute the halves
__device__ __forceinline__ uint32_t flags_to_mask16b(uint32_t f0, uint32_t f1, uint32_t f2, uint32_t f3) {
    const uint32_t h0 = __umulhi(f0, 0x02040810u), h1 = __umulhi(f1, 0x20408100u);
    const uint32_t h2 = __umulhi(f2, 0x02040810u), h3 = __umulhi(f3, 0x20408100u);
    const uint32_t lo = (h0 & 0x0fu) | (h1 & ~0x0fu);  // byte 0 valid
    const uint32_t hi = (h2 & 0x0fu) | (h3 & ~0x0fu);
    return __byte_perm(lo, hi, 0x7740);  // byte0(lo), byte0(hi), zero... selector 7 = byte 3 of hi: cleared below
}

// shared-memory layout of the lean kernel: tile ring, two masks, barriers, then the GROUP BY area
template <class G>
struct LeanLayout {
    static constexpr int OFF_TM = G::STAGES * G::BUF;
    static constexpr int OFF_DM = OFF_TM + G::MASKW * 4;
    static constexpr int OFF_MBAR = OFF_DM + G::MASKW * 4;
    static constexpr int OFF_TABLE = (OFF_MBAR + G::STAGES * 8 + 127) / 128 * 128;
};

// ONELEAF: the commonest shape, `COUNT(*) ... WHERE column <op> decimal literal` (one wanted field, no
// aggregate state): the WHERE program loop, slot selects and operand bookkeeping compile away.
template <class G, int MINB, bool GROUPED, bool ONELEAF, bool MINMAX>
__global__ void __launch_bounds__(G::THREADS, MINB) lean_kernel(const __grid_constant__ DevPlan P) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    using LL = LeanLayout<G>;
    constexpr int kLeanAggBlock = lean_agg_block(MINMAX), kLeanWarpAcc = lean_warp_acc(MINMAX);
    const uint32_t s_tm = sbase + LL::OFF_TM, s_dm = sbase + LL::OFF_DM;
    uint64_t* mbar = (uint64_t*)(smem + LL::OFF_MBAR);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t* dict = smem + LL::OFF_TABLE;                              // GROUPED only
    uint8_t* wacc = dict + kLeanDictCap * kLeanDictEntry + 16 + warp * kLeanWarpAcc;
    unsigned int* ngroups = (unsigned int*)(dict + kLeanDictCap * kLeanDictEntry);

    if (tid == 0) {
        for (int s = 0; s < G::STAGES; s++) mbar_init(&mbar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int w = G::BUF / 32 + tid; w < G::MASKW; w += G::THREADS) {
        sts32(s_tm + 4 * w, 0xffffffffu);
        sts32(s_dm + 4 * w, 0u);
    }
    if (GROUPED) {
        const int words = (kLeanDictCap * kLeanDictEntry + 16 + G::NWARPS * kLeanWarpAcc) / 4;
        for (int k = tid; k < words; k += G::THREADS) ((uint32_t*)dict)[k] = 0u;
        __syncthreads();
        for (int k = tid; k < kLeanDictCap; k += G::THREADS) *(uint64_t*)(dict + k * kLeanDictEntry + 16) = ~0ull;  // first okey
        for (int a = 0; a < 4; a++) {
            if (MINMAX && a < P.l_nagg && (P.aggs[P.l_agg[a]].func == CQG_AGG_MIN || P.aggs[P.l_agg[a]].func == CQG_AGG_MAX)) {
                const uint64_t empty = P.aggs[P.l_agg[a]].func == CQG_AGG_MIN ? ~0ull : 0ull;
                for (int k = tid; k < G::NWARPS * kLeanGroups; k += G::THREADS) {
                    uint64_t* st = (uint64_t*)(dict + kLeanDictCap * kLeanDictEntry + 16 + (k / kLeanGroups) * kLeanWarpAcc +
                                               kLeanGroups * 4 + a * kLeanAggBlock + 32 * (k % kLeanGroups));
                    st[0] = ~0ull;
                    st[2] = empty;
                    st[3] = ~0ull;
                }
            }
        }
    }
    __syncthreads();

    uint32_t rows = 0, count = 0;
    uint64_t first = ~0ull;
    bool have_first = false;
    long long s3[4];
    uint32_t sn[4];
    uint64_t mmk[4], mmo[4], mmf[4];  // scalar MIN/MAX (MINMAX && !GROUPED): extreme key, its row, first non-NULL row
#pragma unroll
    for (int a = 0; a < 4; a++) {
        s3[a] = 0;
        sn[a] = 0;
        mmk[a] = 0;
        mmo[a] = ~0ull;
        mmf[a] = ~0ull;
    }
    // plan constants the row loop uses, once
    const uint64_t size = P.size;
    const uint32_t patD = (uint32_t)P.delim * 0x01010101u;
    uint32_t patN, patDv;  // opaque to the compiler: stay in vector registers
    uint32_t one;
    asm volatile("mov.u32 %0, 1;" : "=r"(one));
    asm volatile("mov.u32 %0, 0x0a0a0a0a;" : "=r"(patN));
    asm volatile("mov.u32 %0, %1;" : "=r"(patDv) : "r"(patD));
    const int nwant = CQG_SPEC(NWANT, P.nwantL);
    const int gap0 = CQG_SPEC(GAP0, P.gap[0]), gap1 = CQG_SPEC(GAP1, P.gap[1]), gap2 = CQG_SPEC(GAP2, P.gap[2]),
              gap3 = CQG_SPEC(GAP3, P.gap[3]);
    const int nprog = CQG_SPEC(NPROG, P.l_nprog);
    const int nagg = CQG_SPEC(NAGG, P.l_nagg);
    const int leaf0_lop = P.l_leaf[0].lop;
    const int ngc = GROUPED ? CQG_SPEC(NGC, P.ngc) : 0;
    uint32_t summask = 0;  // aggregates that read a column: SUM/AVG, and (bits 4..7) those that are MIN/MAX, (8..11) MIN
    int aslot[4];
#pragma unroll
    for (int a = 0; a < 4; a++) {
        aslot[a] = 0;
        if (a < nagg) {
            const AggSpec sp = P.aggs[P.l_agg[a]];
            const int func = CQG_SPEC_AT(AFUNC, a, sp.func);
            summask |= 1u << a;
            if (func == CQG_AGG_MIN || func == CQG_AGG_MAX) summask |= 16u << a;
            if (func == CQG_AGG_MIN) summask |= 256u << a;
            aslot[a] = CQG_SPEC_AT(ASLOT, a, sp.slot);
            if (func == CQG_AGG_MIN) mmk[a] = ~0ull;
        }
    }

    auto issue = [&](int it) {
        const long long tile = (long long)P.first_tile + blockIdx.x + (long long)it * gridDim.x;
        const int stage = it % G::STAGES;
        const long long g0 = tile * (long long)G::TILE - G::PRE;
        if (g0 >= 0 && g0 + G::BUF <= (long long)size) {  // edge tiles are handed over, not loaded
            mbar_expect_tx(&mbar[stage], G::BUF);
            tma_load_1d(smem + G::OFF_BUF + stage * G::BUF, P.data + g0, G::BUF, &mbar[stage]);
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
        const unsigned abort_now = (*(volatile unsigned*)P.errflags) & KERR_LEAN_ABORT;
        mbar_wait(&mbar[stage], parity);
        const uint32_t s_buf = sbase + G::OFF_BUF + stage * G::BUF;
        const long long tile = (long long)P.first_tile + blockIdx.x + (long long)it * gridDim.x;
        const long long g0 = tile * (long long)G::TILE - G::PRE;
        const bool edge = g0 < 0 || g0 + G::BUF > (long long)size;

        // ---- phase 1: terminator / delimiter masks, and "is the tile clean" ----
        uint32_t spec = edge ? 0x80u : 0u;
        if (!edge) {
#pragma unroll 2
            for (int c = tid; c < G::CHUNKS; c += G::THREADS) {
                const uint4 v = lds128(s_buf + 16 * c);
                // 0x80 where the byte equals the pattern: ((v ^ pat) & 0x7f..) + 0x7f.. has bit 7 set iff the low 7
                // bits differ; OR v brings in bit 7 of the byte itself. Patterns sit in registers so that the first
                // step is ONE lop3 with the 0x7f.. immediate.
                const uint32_t f0 = ~(add_fma((v.x ^ patN) & 0x7f7f7f7fu, one, 0x7f7f7f7fu) | v.x) & 0x80808080u;
                const uint32_t f1 = ~(add_fma((v.y ^ patN) & 0x7f7f7f7fu, one, 0x7f7f7f7fu) | v.y) & 0x80808080u;
                const uint32_t f2 = ~(add_fma((v.z ^ patN) & 0x7f7f7f7fu, one, 0x7f7f7f7fu) | v.z) & 0x80808080u;
                const uint32_t f3 = ~(add_fma((v.w ^ patN) & 0x7f7f7f7fu, one, 0x7f7f7f7fu) | v.w) & 0x80808080u;
                const uint32_t d0 = ~(add_fma((v.x ^ patDv) & 0x7f7f7f7fu, one, 0x7f7f7f7fu) | v.x) & 0x80808080u;
                const uint32_t d1 = ~(add_fma((v.y ^ patDv) & 0x7f7f7f7fu, one, 0x7f7f7f7fu) | v.y) & 0x80808080u;
                const uint32_t d2 = ~(add_fma((v.z ^ patDv) & 0x7f7f7f7fu, one, 0x7f7f7f7fu) | v.z) & 0x80808080u;
                const uint32_t d3 = ~(add_fma((v.w ^ patDv) & 0x7f7f7f7fu, one, 0x7f7f7f7fu) | v.w) & 0x80808080u;
                // "special" bytes (kLeanSpecialXor): controls and the quote, not the blank; '\n' is lifted out of the way first
                const uint32_t x0 = (v.x | f0) ^ kLeanSpecialXor, x1 = (v.y | f1) ^ kLeanSpecialXor, x2 = (v.z | f2) ^ kLeanSpecialXor,
                               x3 = (v.w | f3) ^ kLeanSpecialXor;
                spec |= (add_fma(x0, one, kLeanSpecialSub) & ~x0) | (add_fma(x1, one, kLeanSpecialSub) & ~x1) |
                        (add_fma(x2, one, kLeanSpecialSub) & ~x2) | (add_fma(x3, one, kLeanSpecialSub) & ~x3);  // x - 0x21212121
                sts16(s_tm + 2 * c, flags_to_mask16b(f0, f1, f2, f3));
                sts16(s_dm + 2 * c, flags_to_mask16b(d0, d1, d2, d3));
            }
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

        // ---- phase 2: every thread walks the rows that START in its own 32*WPT bytes ----
        long long olo_l = (long long)P.own_lo - g0, ohi_l = (long long)P.own_hi - g0;
        const uint32_t olo = (uint32_t)(olo_l < G::PRE ? G::PRE : (olo_l > G::PRE + G::TILE ? G::PRE + G::TILE : olo_l));
        const uint32_t ohi = (uint32_t)(ohi_l < G::PRE ? G::PRE : (ohi_l > G::PRE + G::TILE ? G::PRE + G::TILE : ohi_l));
        const bool partial_tile = olo > (uint32_t)G::PRE || ohi < (uint32_t)(G::PRE + G::TILE);
        const uint32_t w0 = G::PRE / 32 + tid * G::WPT;
        uint32_t handed = 0, myrows = 0;
        static_assert(G::WPT == 4, "the row walk below keeps four start words in registers");
        uint32_t S0, S1, S2, S3;
        {
            uint32_t prev = lds32(s_tm + 4 * (w0 - 1));
            uint32_t S[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t t = lds32(s_tm + 4 * (w0 + j));
                uint32_t s = ((t << 1) | (prev >> 31)) & ~t;  // row starts: a terminator, then a byte that is none
                prev = t;
                if (partial_tile) {  // only the first and last tile of a shard own less than all of their bytes
                    const uint32_t p0 = (w0 + j) * 32u;
                    if (olo > p0) s &= olo >= p0 + 32u ? 0u : (0xffffffffu << (olo - p0));
                    if (ohi < p0 + 32u) s &= ohi <= p0 ? 0u : (0xffffffffu >> (p0 + 32u - ohi));
                }
                S[j] = s;
            }
            S0 = S[0];
            S1 = S[1];
            S2 = S[2];
            S3 = S[3];
        }
        // all rows of the span back to back, so that lanes with 4 and lanes with 5 rows stay together
        {
            int j = 0;
            uint32_t s = S0;
            for (;;) {
                while (s == 0u && j < 3) {
                    j++;
                    s = j == 1 ? S1 : (j == 2 ? S2 : S3);
                }
                const bool has = s != 0u;
                if (!has) break;
                bool ok = true, pass = true;
                uint32_t gid = 0xffffffffu, addmask = 0, rs = 0;
                unsigned long long add0 = 0, add1 = 0, add2 = 0, add3 = 0;
                if (has) {
                    const uint32_t bi = __ffs(s) - 1u;
                    s &= s - 1u;
                    const uint32_t t = lds32(s_tm + 4 * (w0 + j));
                    rs = (w0 + j) * 32u + bi;
                    const uint32_t wa = 4u * (w0 + j);
                    myrows++;
                    uint32_t off0 = 0, off1 = 0, off2 = 0, off3 = 0, len0 = 0, len1 = 0, len2 = 0, len3 = 0;
                    const uint32_t t1 = lds32(s_tm + wa + 4);
                    const uint32_t tw = __funnelshift_r(t, t1, bi);
                    if (tw) {
                        // the row ends inside a 32-bit window
                        const uint32_t len = __ffs(tw) - 1u;
                        uint32_t dw = __funnelshift_r(lds32(s_dm + wa), lds32(s_dm + wa + 4), bi) & ((1u << len) - 1u);
                        uint32_t sp = 0;
                        bool missing = false;
#define CQG_LEAN_FIELD(K, GAP, OFF, LEN)                                 \
    if (nwant > K) {                                                     \
        if (GAP > 0) {                                                   \
            for (int i = 1; i < GAP; i++) dw &= dw - 1u;                 \
            missing = missing || dw == 0u;                               \
            sp = __ffs(dw);                                              \
            dw &= dw - 1u;                                               \
        }                                                                \
        const uint32_t ep = dw ? __ffs(dw) - 1u : len;                   \
        OFF = rs + sp;                                                   \
        LEN = missing ? 0u : ep - sp;                                    \
    }
                        CQG_LEAN_FIELD(0, gap0, off0, len0)
                        CQG_LEAN_FIELD(1, gap1, off1, len1)
                        CQG_LEAN_FIELD(2, gap2, off2, len2)
                        CQG_LEAN_FIELD(3, gap3, off3, len3)
#undef CQG_LEAN_FIELD
                    } else {
                        const uint32_t t2 = lds32(s_tm + wa + 8);
                        const uint32_t tw2 = __funnelshift_r(t1, t2, bi);
                        if (tw2 == 0u) {
                            ok = false;  // 64 bytes or more
                        } else {
                            const uint32_t len = 32u + __ffs(tw2) - 1u;
                            const uint32_t d0 = lds32(s_dm + wa), d1 = lds32(s_dm + wa + 4), d2 = lds32(s_dm + wa + 8);
                            unsigned long long dw =
                                (((unsigned long long)__funnelshift_r(d1, d2, bi) << 32) | __funnelshift_r(d0, d1, bi)) &
                                ((1ull << len) - 1ull);
                            uint32_t sp = 0;
                            bool missing = false;
#define CQG_LEAN_FIELD(K, GAP, OFF, LEN)                                 \
    if (nwant > K) {                                                     \
        if (GAP > 0) {                                                   \
            for (int i = 1; i < GAP; i++) dw &= dw - 1ull;               \
            missing = missing || dw == 0ull;                             \
            sp = (uint32_t)__ffsll((long long)dw);                       \
            dw &= dw - 1ull;                                             \
        }                                                                \
        const uint32_t ep = dw ? (uint32_t)__ffsll((long long)dw) - 1u : len; \
        OFF = rs + sp;                                                   \
        LEN = missing ? 0u : ep - sp;                                    \
    }
                            CQG_LEAN_FIELD(0, gap0, off0, len0)
                            CQG_LEAN_FIELD(1, gap1, off1, len1)
                            CQG_LEAN_FIELD(2, gap2, off2, len2)
                            CQG_LEAN_FIELD(3, gap3, off3, len3)
#undef CQG_LEAN_FIELD
                        }
                    }
#define CQG_LEAN_SLOT(SL, O, L)                                                  \
    const uint32_t O = SL == 0 ? off0 : SL == 1 ? off1 : SL == 2 ? off2 : off3; \
    const uint32_t L = SL == 0 ? len0 : SL == 1 ? len1 : SL == 2 ? len2 : len3;
                    // ---- WHERE: leaves on short decimals and short texts, combined on a bit stack ----
                    if (ONELEAF) {
                        uint32_t mant, fd;
                        bool hd;
                        if (ok && len0 - 1u < 7u && lean_decimal(s_buf + off0, len0, mant, fd, hd)) {
                            const long long lhs = (long long)((unsigned long long)mant * (unsigned long long)P.l_leaf[0].A[fd]);
                            const long long rhs = P.l_leaf[0].LB[fd];
                            pass = leaf0_lop == 0 ? lhs > rhs : leaf0_lop == 1 ? lhs < rhs : leaf0_lop == 2 ? lhs == rhs : lhs != rhs;
                        } else {
                            ok = false;
                        }
                    } else if (ok && nprog) {
                        uint32_t bs = 0;
                        CQG_SPEC_UNROLL
                        for (int pc = 0; pc < nprog; pc++) {
                            const int c = CQG_SPEC_AT(PROG, pc, P.l_prog[pc]);
                            if (c >= 0) {
                                const int sl = CQG_SPEC_AT(LEAFSLOT, c, P.l_leaf[c].slot), kind = CQG_SPEC_AT(LEAFKIND, c, P.l_leaf[c].kind);
                                CQG_LEAN_SLOT(sl, o, l)
                                bool bv = false;
                                if (kind == 0) {
                                    uint32_t mant, fd;
                                    bool hd;
                                    if (l - 1u < 7u && lean_decimal(s_buf + o, l, mant, fd, hd)) {
                                        const long long lhs = (long long)((unsigned long long)mant * (unsigned long long)P.l_leaf[c].A[fd]);
                                        const long long rhs = P.l_leaf[c].LB[fd];
                                        const int lop = P.l_leaf[c].lop;
                                        bv = lop == 0 ? lhs > rhs : lop == 1 ? lhs < rhs : lop == 2 ? lhs == rhs : lhs != rhs;
                                    } else {
                                        ok = false;  // NULL, text, date, signed or long number: general kernel
                                    }
                                } else {
                                    // text equality: value_compare is strcmp for two strings, "less" for a NULL field,
                                    // and 0 ("equal") for a number or date against text - those rows are handed over
                                    uint32_t tag;
                                    uint64_t w0, w1;
                                    if (l == 0u) {
                                        bv = kind == 2;
                                    } else if (l > 16u) {
                                        const uint32_t c0 = lds8(s_buf + o);
                                        const bool ns = (c0 - 48u) <= 9u || c0 == '+' || c0 == '-' || c0 == '.';
                                        if (ns || c0 == ' ' || lds8(s_buf + o + l - 1u) == ' ') ok = false;  // (trimmed by the reference)
                                        bv = kind == 2;  // longer than the literal: different
                                    } else if (lean_key_part(s_buf + o, l, tag, w0, w1) && (tag == KT_STR || tag == KT_NULL)) {
                                        // (the text NULL packs as KT_NULL with zero words: compare its bytes directly)
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
                    // ---- SUM / AVG operands ----
                    if (!ONELEAF && ok && pass && (summask & 15u)) {
#define CQG_LEAN_AGG(A, ADD)                                                                             \
    if (summask & (1u << A)) {                                                                           \
        const int sl = aslot[A];                                                                         \
        CQG_LEAN_SLOT(sl, o, l)                                                                          \
        uint32_t mant, fd;                                                                               \
        bool hd;                                                                                         \
        if (l - 1u < 7u && lean_decimal(s_buf + o, l, mant, fd, hd)) {                                   \
            if (MINMAX && (summask & (16u << A))) {                                                      \
                /* MIN/MAX: the key value_compare orders by = the double the reference parses */        \
                const double dv = mant < 10000u ? P.dec_table[fd * 10000u + mant] : (double)mant / kPow10[fd]; \
                ADD = num_key(dv);                                                                       \
            } else {                                                                                     \
                ADD = (unsigned long long)mant * (fd == 0u ? 1000u : fd == 1u ? 100u : fd == 2u ? 10u : 1u); \
            }                                                                                            \
            addmask |= 1u << A;                                                                          \
        } else if (l != 0u) {                                                                            \
            ok = false; /* a value this kernel does not decode (NULL is simply not summed) */           \
        }                                                                                                \
    }
                        CQG_LEAN_AGG(0, add0)
                        CQG_LEAN_AGG(1, add1)
                        CQG_LEAN_AGG(2, add2)
                        CQG_LEAN_AGG(3, add3)
#undef CQG_LEAN_AGG
                    }
                    // ---- GROUP BY: key -> group number of this CTA ----
                    if (GROUPED && ok && pass) {
                        uint64_t kw[8];
                        uint32_t tags = 0;
                        uint64_t h = 0x243F6A8885A308D3ull + (uint64_t)ngc;
#pragma unroll
                        for (int g = 0; g < 4; g++) {
                            kw[2 * g] = 0;
                            kw[2 * g + 1] = 0;
                            if (g < ngc && ok) {
                                const int sl = CQG_SPEC_AT(GSLOT, g, P.gslot[g]);
                                uint32_t tag = KT_NULL;
                                if (sl >= 0) {
                                    CQG_LEAN_SLOT(sl, o, l)
                                    ok = lean_key_part(s_buf + o, l, tag, kw[2 * g], kw[2 * g + 1]);
                                }
                                tags |= tag << (4 * g);
                                h = key_hash_step(h, tag, kw[2 * g], kw[2 * g + 1]);
                            }
                        }
                        if (ok) {
                            h = key_hash_final(h);
                            // find-or-insert in the CTA dictionary
                            uint32_t i = (uint32_t)(h >> 1) & (kLeanDictCap - 1);
                            for (int probes = 0; probes < kLeanDictCap;) {
                                uint8_t* e = dict + i * kLeanDictEntry;
                                unsigned long long cur = *(volatile unsigned long long*)e;
                                if (cur == 0ull) {
                                    cur = atomicCAS((unsigned long long*)e, 0ull, (unsigned long long)(h | kLockBit));
                                    if (cur == 0ull) {
                                        const unsigned int id = atomicAdd(ngroups, 1u);
                                        *(uint32_t*)(e + 8) = id;
                                        *(uint32_t*)(e + 12) = tags;
#pragma unroll
                                        for (int g = 0; g < 4; g++) {
                                            *(uint64_t*)(e + 32 + 16 * g) = kw[2 * g];
                                            *(uint64_t*)(e + 40 + 16 * g) = kw[2 * g + 1];
                                        }
                                        __threadfence_block();
                                        atomicExch((unsigned long long*)e, (unsigned long long)h);
                                        gid = id;
                                        break;
                                    }
                                }
                                if ((cur & ~kLockBit) == h) {
                                    if (cur & kLockBit) continue;  // being written: look again
                                    bool same = *(volatile uint32_t*)(e + 12) == tags;
#pragma unroll
                                    for (int g = 0; g < 4; g++)
                                        same = same && *(volatile uint64_t*)(e + 32 + 16 * g) == kw[2 * g] &&
                                               *(volatile uint64_t*)(e + 40 + 16 * g) == kw[2 * g + 1];
                                    if (same) {
                                        gid = *(volatile uint32_t*)(e + 8);
                                        break;
                                    }
                                }
                                i = (i + 1) & (kLeanDictCap - 1);
                                probes++;
                            }
                            if (gid >= (uint32_t)kLeanGroups) {
                                // more groups than a CTA numbers: rerun on the global table
                                atomicOr(P.errflags, KERR_LEAN_ABORT | KERR_LEAN_GROUPS);
                                gid = 0xffffffffu;
                            } else {
                                const uint64_t okey = (P.global_base + (uint64_t)(g0 + (long long)rs)) << 16;
                                uint64_t* fp = (uint64_t*)(dict + i * kLeanDictEntry + 16);
                                if (okey < *(volatile uint64_t*)fp) atomicMin((unsigned long long*)fp, (unsigned long long)okey);
                            }
                        }
                    }
#undef CQG_LEAN_SLOT
                    if (!ok) {
                        unsigned long long k = atomicAdd(P.def_row_count, 1ull);
                        if (k < P.def_row_cap) P.def_rows[k] = (uint64_t)(g0 + (long long)rs);
                        handed++;
                        gid = 0xffffffffu;
                    } else {
                        rows++;
                    }
                }
                const bool take = has && ok && pass;
                if (!GROUPED) {
                    if (take) {
                        count++;
                        const uint64_t gabs = P.global_base + (uint64_t)(g0 + (long long)rs);
                        // a thread meets its rows in increasing offset order (within a tile and from tile to tile)
                        if (!have_first) {
                            first = gabs;
                            have_first = true;
                        }
                        if (!ONELEAF && (addmask & 1u)) {
                            if (MINMAX && (summask & (16u << 0))) {
                                const uint64_t ok_ = gabs << 16;
                                const bool mn = (summask & (256u << 0)) != 0u;
                                if (mn ? (add0 < mmk[0] || (add0 == mmk[0] && ok_ < mmo[0])) : (add0 > mmk[0] || (add0 == mmk[0] && ok_ < mmo[0]))) {
                                    mmk[0] = add0;
                                    mmo[0] = ok_;
                                }
                                if (ok_ < mmf[0]) mmf[0] = ok_;
                            } else {
                                s3[0] += (long long)add0;
                                sn[0]++;
                            }
                        }
                        if (!ONELEAF && (addmask & 2u)) {
                            if (MINMAX && (summask & (16u << 1))) {
                                const uint64_t ok_ = gabs << 16;
                                const bool mn = (summask & (256u << 1)) != 0u;
                                if (mn ? (add1 < mmk[1] || (add1 == mmk[1] && ok_ < mmo[1])) : (add1 > mmk[1] || (add1 == mmk[1] && ok_ < mmo[1]))) {
                                    mmk[1] = add1;
                                    mmo[1] = ok_;
                                }
                                if (ok_ < mmf[1]) mmf[1] = ok_;
                            } else {
                                s3[1] += (long long)add1;
                                sn[1]++;
                            }
                        }
                        if (!ONELEAF && (addmask & 4u)) {
                            if (MINMAX && (summask & (16u << 2))) {
                                const uint64_t ok_ = gabs << 16;
                                const bool mn = (summask & (256u << 2)) != 0u;
                                if (mn ? (add2 < mmk[2] || (add2 == mmk[2] && ok_ < mmo[2])) : (add2 > mmk[2] || (add2 == mmk[2] && ok_ < mmo[2]))) {
                                    mmk[2] = add2;
                                    mmo[2] = ok_;
                                }
                                if (ok_ < mmf[2]) mmf[2] = ok_;
                            } else {
                                s3[2] += (long long)add2;
                                sn[2]++;
                            }
                        }
                        if (!ONELEAF && (addmask & 8u)) {
                            if (MINMAX && (summask & (16u << 3))) {
                                const uint64_t ok_ = gabs << 16;
                                const bool mn = (summask & (256u << 3)) != 0u;
                                if (mn ? (add3 < mmk[3] || (add3 == mmk[3] && ok_ < mmo[3])) : (add3 > mmk[3] || (add3 == mmk[3] && ok_ < mmo[3]))) {
                                    mmk[3] = add3;
                                    mmo[3] = ok_;
                                }
                                if (ok_ < mmf[3]) mmf[3] = ok_;
                            } else {
                                s3[3] += (long long)add3;
                                sn[3]++;
                            }
                        }
                    }
                } else {
                    // ---- group step: this warp's own accumulators in shared memory, native 32-bit
                    // atomics only; a 64-bit sum is (lo, hi) with the carry added by the lane that wrapped lo ----
                    const uint64_t okey_row = (P.global_base + (uint64_t)(g0 + (long long)rs)) << 16;
                    if (take && gid != 0xffffffffu) {
                        atomicAdd((unsigned int*)(wacc + 4 * gid), 1u);
#define CQG_LEAN_GSUM(A, ADD)                                                                  \
    if ((addmask >> A) & 1u) {                                                                 \
        uint8_t* b = wacc + kLeanGroups * 4 + A * kLeanAggBlock;                               \
        if (MINMAX && (summask & (16u << A))) {                                                \
            uint64_t* st = (uint64_t*)(b + 32 * gid);                                          \
            amin64(&st[0], (okey_row << 2) | 1u);                                              \
            num_extreme(&st[2], ADD, okey_row, (summask & (256u << A)) != 0u);                 \
        } else {                                                                               \
            const uint32_t vlo = (uint32_t)ADD, vhi = (uint32_t)(ADD >> 32);                   \
            const uint32_t old = atomicAdd((unsigned int*)(b + 4 * gid), vlo);                 \
            const uint32_t up = vhi + ((old + vlo) < old ? 1u : 0u);                           \
            if (up) atomicAdd((unsigned int*)(b + kLeanGroups * 4 + 4 * gid), up);             \
            atomicAdd((unsigned int*)(b + kLeanGroups * 8 + 4 * gid), 1u);                     \
        }                                                                                      \
    }
                        CQG_LEAN_GSUM(0, add0)
                        CQG_LEAN_GSUM(1, add1)
                        CQG_LEAN_GSUM(2, add2)
                        CQG_LEAN_GSUM(3, add3)
#undef CQG_LEAN_GSUM
                    }
                }
            }
        }
        // too many rows outside this kernel's repertoire: let the general kernel do the whole scan.
        // The barrier also keeps the tile and its masks alive until every thread is done with them.
        const int many = __syncthreads_or((int)(handed * 8u > myrows + 8u));
        if (many && tid == 0) atomicOr(P.errflags, KERR_LEAN_ABORT);
    }

    // ---- epilogue ----
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) rows += __shfl_xor_sync(0xffffffffu, rows, d);
    if (lane == 0 && rows) atomicAdd(P.rows_scanned, (unsigned long long)rows);
    if (!GROUPED) {
        // fold the registers into the single group `_all_`
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            count += __shfl_xor_sync(0xffffffffu, count, d);
            const uint64_t of = __shfl_xor_sync(0xffffffffu, first, d);
            first = of < first ? of : first;
#pragma unroll
            for (int a = 0; a < 4; a++) {
                s3[a] += __shfl_xor_sync(0xffffffffu, s3[a], d);
                sn[a] += __shfl_xor_sync(0xffffffffu, sn[a], d);
            }
        }
        if (MINMAX) {
            // warp-reduce the extremes (lexicographic on (key, okey)), then lane 0 merges them below
#pragma unroll
            for (int a = 0; a < 4; a++) {
                const bool mn = (summask & (256u << a)) != 0u;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    const uint64_t k2 = __shfl_xor_sync(0xffffffffu, mmk[a], d), o2 = __shfl_xor_sync(0xffffffffu, mmo[a], d);
                    const uint64_t f2 = __shfl_xor_sync(0xffffffffu, mmf[a], d);
                    if (mn ? (k2 < mmk[a] || (k2 == mmk[a] && o2 < mmo[a])) : (k2 > mmk[a] || (k2 == mmk[a] && o2 < mmo[a]))) {
                        mmk[a] = k2;
                        mmo[a] = o2;
                    }
                    if (f2 < mmf[a]) mmf[a] = f2;
                }
            }
        }
        if (lane == 0 && count) {
            unsigned err = 0;
            const uint64_t h = key_hash_final(0x243F6A8885A308D3ull);
            uint8_t* ge = global_entry_for(P, h, 0u, nullptr, err);
            if (ge) {
                atomicAdd((unsigned long long*)(ge + kOffCount), (unsigned long long)count);
                amin64((uint64_t*)(ge + kOffFirst), first << 16);
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    if (MINMAX && a < P.l_nagg && (summask & (16u << a))) {
                        if (mmf[a] != ~0ull) {
                            uint64_t* st = (uint64_t*)(ge + P.aggs[P.l_agg[a]].off);
                            amin64(&st[0], (mmf[a] << 2) | 1u);
                            num_extreme(&st[2], mmk[a], mmo[a], (summask & (256u << a)) != 0u);
                        }
                    } else if (a < P.l_nagg && sn[a]) {
                        atomicAdd((unsigned long long*)(ge + P.aggs[P.l_agg[a]].off + 16), (unsigned long long)sn[a]);
                        atomicAdd((unsigned long long*)(ge + P.aggs[P.l_agg[a]].off + 24), (unsigned long long)s3[a]);
                    }
                }
            }
            if (err) atomicOr(P.errflags, err);
        }
    } else {
        // every dictionary entry: add up the warps' accumulators and fold them into the global table
        __syncthreads();
        unsigned err = 0;
        for (int sidx = tid; sidx < kLeanDictCap; sidx += G::THREADS) {
            const uint8_t* e = dict + sidx * kLeanDictEntry;
            const uint64_t h = *(const uint64_t*)e;
            if (h == 0ull) continue;
            const uint32_t gid = *(const uint32_t*)(e + 8);
            if (gid >= (uint32_t)kLeanGroups) continue;
            unsigned long long c = 0;
            for (int w = 0; w < G::NWARPS; w++) c += *(const uint32_t*)(dict + kLeanDictCap * kLeanDictEntry + 16 + w * kLeanWarpAcc + 4 * gid);
            if (c == 0ull) continue;
            uint64_t kw[2 * CQG_MAX_GROUP_COLS];
            for (int g = 0; g < 4; g++) {
                kw[2 * g] = *(const uint64_t*)(e + 32 + 16 * g);
                kw[2 * g + 1] = *(const uint64_t*)(e + 40 + 16 * g);
            }
            uint8_t* ge = global_entry_for(P, h, *(const uint32_t*)(e + 12), kw, err);
            if (!ge) continue;
            atomicAdd((unsigned long long*)(ge + kOffCount), c);
            amin64((uint64_t*)(ge + kOffFirst), *(const uint64_t*)(e + 16));
            for (int a = 0; a < 4; a++) {
                if (MINMAX && a < P.l_nagg && (P.aggs[P.l_agg[a]].func == CQG_AGG_MIN || P.aggs[P.l_agg[a]].func == CQG_AGG_MAX)) {
                    const bool is_min = P.aggs[P.l_agg[a]].func == CQG_AGG_MIN;
                    uint64_t* gs = (uint64_t*)(ge + P.aggs[P.l_agg[a]].off);
                    for (int w = 0; w < G::NWARPS; w++) {
                        const uint64_t* st = (const uint64_t*)(dict + kLeanDictCap * kLeanDictEntry + 16 + w * kLeanWarpAcc +
                                                               kLeanGroups * 4 + a * kLeanAggBlock + 32 * gid);
                        if (st[0] == ~0ull) continue;
                        amin64(&gs[0], st[0]);
                        num_extreme(&gs[2], st[2], st[3], is_min);
                    }
                } else if (a < P.l_nagg) {
                    unsigned long long t3 = 0, tn = 0;
                    for (int w = 0; w < G::NWARPS; w++) {
                        const uint8_t* b = dict + kLeanDictCap * kLeanDictEntry + 16 + w * kLeanWarpAcc + kLeanGroups * 4 + a * kLeanAggBlock;
                        t3 += ((unsigned long long)*(const uint32_t*)(b + kLeanGroups * 4 + 4 * gid) << 32) + *(const uint32_t*)(b + 4 * gid);
                        tn += *(const uint32_t*)(b + kLeanGroups * 8 + 4 * gid);
                    }
                    if (tn) {
                        atomicAdd((unsigned long long*)(ge + P.aggs[P.l_agg[a]].off + 16), tn);
                        atomicAdd((unsigned long long*)(ge + P.aggs[P.l_agg[a]].off + 24), t3);
                    }
                }
            }
        }
        if (err) atomicOr(P.errflags, err);
    }
}

#ifndef CQG_JIT
// rows handed over one by one: each is found and split straight from HBM by the general operators
__global__ void deferred_rows_kernel(const __grid_constant__ DevPlan P, const uint64_t* rows, uint64_t n) {
    CtaState cs;
    cs.stab = nullptr;
    cs.s_occ = nullptr;
    ThreadAcc acc;
    acc.rows = 0;
    acc.count = 0;
    acc.first = ~0ull;
    acc.err = 0;
    for (int a = 0; a < 4; a++) {
        acc.si[a] = 0;
        acc.sd[a] = 0.0;
        acc.s3[a] = 0;
        acc.sn[a] = 0;
    }
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        acc.rows++;
        process_long_row(P, cs, rows[i], acc);
    }
    if (acc.rows) atomicAdd(P.rows_scanned, (unsigned long long)acc.rows);
    if (acc.err) atomicOr(P.errflags, acc.err);
}

// RIGHT / FULL joins: the right rows without a match (emit_right_only_row), one thread per slot of the join table
__global__ void join_unmatched_right_kernel(const __grid_constant__ DevPlan P) {
    CtaState cs;
    cs.stab = nullptr;
    cs.s_occ = nullptr;
    ThreadAcc acc;
    acc.rows = 0;
    acc.count = 0;
    acc.first = ~0ull;
    acc.err = 0;
    for (int a = 0; a < 4; a++) {
        acc.si[a] = 0;
        acc.sd[a] = 0.0;
        acc.s3[a] = 0;
        acc.sn[a] = 0;
    }
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < P.jcap; i += (uint64_t)gridDim.x * blockDim.x) {
        const JoinSlot* s = &P.jslots[i];
        if (s->h == 0ull || (s->tag & kJoinMatched)) continue;
        for (uint32_t it = s->head; it != 0u; it = P.jrow_next[it - 1]) emit_right_only_row(P, cs, P.jrow_off[it - 1], acc);
    }
    if (acc.err) atomicOr(P.errflags, acc.err);
}

// ---- packed table -> general entries ----
// Every occupied line of the packed table becomes one general entry (DevPlan::entry_bytes, the image the host
// finish path, the partial export and merge_entries_kernel read), written densely in arbitrary order; the
// entry's index is left in the line's count word for merge_general_into_dense_kernel.
__global__ void expand_packed_kernel(const __grid_constant__ DevPlan P, uint8_t* out, uint64_t out_cap, unsigned long long* n_out) {
    const int eb = P.entry_bytes, pb = P.pk.entry_bytes;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < P.pcap; i += (uint64_t)gridDim.x * blockDim.x) {
        uint8_t* e = P.ptab + i * (uint64_t)pb;
        if (*(const uint64_t*)e == 0ull) continue;
        const unsigned long long idx = atomicAdd(n_out, 1ull);
        if (idx >= out_cap) continue;
        uint8_t* o = out + idx * (uint64_t)eb;
        copy_entry_init(o, P.entry_init, eb);
        const uint64_t fw = *(const uint64_t*)e, count = *(const uint64_t*)(e + P.pk.count_off);
        const uint64_t first = (fw & ~0xffffull) - kPkOkeyBias;
        const uint32_t tags = (uint32_t)(fw & 0xffffull);
        *(uint64_t*)(e + P.pk.count_off) = idx;
        *(uint64_t*)(o + kOffFirst) = first;
        *(uint64_t*)(o + kOffCount) = count;
        *(uint32_t*)(o + kOffTags) = tags;
        uint64_t h = 0x243F6A8885A308D3ull + (uint64_t)P.ngc;  // the general table's hash (agg_row, cqg_scan.cuh), once per group
        uint64_t kw0[4] = {0, 0, 0, 0};
        for (int g = 0; g < P.ngc && g < 4; g++) {
            const uint64_t w0 = *(const uint64_t*)(e + 8 * P.pk.key_word[g]) ^ kPkKeyMask;
            const uint64_t w1 = P.pk.key_wide[g] ? *(const uint64_t*)(e + 8 * P.pk.key_word[g] + 8) ^ kPkKeyMask : 0ull;
            kw0[g] = w0;
            *(uint64_t*)(o + kOffKeys + 16 * g) = w0;
            *(uint64_t*)(o + kOffKeys + 16 * g + 8) = w1;
            h = key_hash_step(h, (tags >> (4 * g)) & 15u, w0, w1);
        }
        *(uint64_t*)(o + kOffHash) = key_hash_final(h);
        for (int a = 0; a < P.l_nagg; a++) {
            const AggSpec sp = P.aggs[P.l_agg[a]];
            uint64_t* st = (uint64_t*)(o + sp.off);
            const bool mm = sp.func == CQG_AGG_MIN || sp.func == CQG_AGG_MAX;
            const int kg = P.pk.agg_key[a];
            if (kg >= 0) {
                // the aggregate's column is GROUP BY column kg: one value per group (INTEGER w0, or DOUBLE w0 / 10^6
                // with w0 a multiple of 1000: rows whose part is anything else were handed over)
                const bool is_int = ((tags >> (4 * kg)) & 15u) == KT_INT;
                if (mm) {
                    const double dv = is_int ? (double)kw0[kg] : __ddiv_rn((double)kw0[kg], 1000000.0);
                    st[0] = (first << 2) | 1u;
                    st[2] = num_key(dv);
                    st[3] = first;  // every row of the group ties: the earliest one holds the value
                } else {
                    st[2] = count;
                    st[3] = count * (is_int ? kw0[kg] * 1000ull : kw0[kg] / 1000ull);
                }
            } else {
                const uint64_t* ps = (const uint64_t*)(e + P.pk.agg_off[a]);
                if (mm) {
                    st[0] = (first << 2) | 1u;  // every row of the line had a numeric operand: the earliest one is the group's first
                    st[2] = ps[0];
                    st[3] = ps[1];
                } else {
                    st[2] = count;  // values summed = rows (NULL operands never reach the packed table)
                    st[3] = ps[0];  // sum of value * 1000
                }
            }
        }
    }
}

// Entries the general kernel built for the tiles and rows the lean kernel handed over (`recs`, compacted) are
// folded into the expanded ones: a key the packed table holds is merged into its expanded entry, any other
// is appended. Two records never share a key (they come out of one hash table).
__global__ void merge_general_into_dense_kernel(const __grid_constant__ DevPlan P, const uint8_t* recs, uint64_t n, uint8_t* dense,
                                                uint64_t dense_cap, unsigned long long* n_dense) {
    const int eb = P.entry_bytes;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint8_t* r = recs + i * (uint64_t)eb;
        const uint32_t tags = *(const uint32_t*)(r + kOffTags);
        uint64_t kw[8];
        bool packable = P.pcap != 0 && P.ngc <= 4;
        for (int g = 0; g < 4; g++) {
            kw[2 * g] = kw[2 * g + 1] = 0;
            if (g < P.ngc) {
                kw[2 * g] = *(const uint64_t*)(r + kOffKeys + 16 * g);
                kw[2 * g + 1] = *(const uint64_t*)(r + kOffKeys + 16 * g + 8);
                // a narrow slot holds w0 only: a key with a second word is not in the packed table
                if (!P.pk.key_wide[g] && kw[2 * g + 1] != 0ull) packable = false;
            }
        }
        uint8_t* pe = nullptr;
        if (packable) {
            uint64_t word0;
            pe = packed_find<false>(P, P.ngc, packed_hash(P.ngc, tags, kw), tags, kw, word0);
        }
        if (pe) {
            entry_merge(P, dense + *(const uint64_t*)(pe + P.pk.count_off) * (uint64_t)eb, r);
        } else {
            const unsigned long long idx = atomicAdd(n_dense, 1ull);
            if (idx < dense_cap) {
                uint8_t* o = dense + idx * (uint64_t)eb;
                for (int k = 0; k < eb; k += 8) *(uint64_t*)(o + k) = *(const uint64_t*)(r + k);
            }
        }
    }
}

#endif  // CQG_JIT

}  // namespace cqg
