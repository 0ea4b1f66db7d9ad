// cqg_lean2.cuh — the scalar lean kernel: DevPlan::simple == 1 plans (no GROUP BY, no MIN/MAX):
// COUNT / SUM / AVG over short decimals, WHERE empty or a postfix program of `column <op> decimal
// literal` / `column = | != 'text'` leaves. Same tile pipeline as cqg_lean.cuh (1-D TMA tile -> SWAR
// masks -> thread-owned row walk) with half the instructions per byte:
//   * two byte classes only: T = "special" (every control, i.e. every terminator, and the quote - not the blank,
//     see kLeanSpecialXor in cqg_lean.cuh) and D = delimiter.
//     Whether a T byte really is '\n' is checked once per ROW (one byte load at the row's end), not
//     once per byte; a tile where any row ends in something else, holds an empty line, or hands over
//     too many rows is DIRTY and goes to the general kernel as a whole (src/csv_reader.c:404-427 row
//     split, :278-338 field split, :195-240 typed decode are then reproduced exactly there);
//   * byte flags become mask bits through dot products (IDP.4A runs at twice the rate of IMAD.HI on sm_100a,
//     tools/micro/pipe_bench.cu), adds are issued as IMAD so that the ALU and FMA pipes share the work;
//   * T and D words interleaved (one 64-bit shared load gives both), a running cursor from row end to
//     next row start (no row-start masks, the next row's mask words asked for before the decode), fields
//     <= 4 bytes decoded right-aligned with one byte-permute that drops the '.';
//   * `mant * A[fd] <op> LB[fd]` (cqg_lean.cuh) is folded, per CTA, into ONE modular interval test
//     (v - lo <= width, != included); the COUNT(*)-WHERE shape compares digit bytes, no multiply at all;
//   * the CTA's next tile is prefetched into L2 while the current one is worked on.
// Results of a tile are committed only after the whole CTA found the tile clean.
#pragma once
#include "cqg_lean.cuh"

namespace cqg {

template <class G>
struct Lean2Layout {
    static constexpr int OFF_MSK = G::STAGES * G::BUF;           // (T word, D word) per 32 CSV bytes
    static constexpr int OFF_CMP = OFF_MSK + G::MASKW * 8;       // kMaxLeanLeaf x 4 fd x {lo, width, code lo, code width}
    static constexpr int OFF_MISC = OFF_CMP + kMaxLeanLeaf * 64; // handed-row counters (2 x u32)
    static constexpr int OFF_MBAR = OFF_MISC + 16;
    static constexpr int TOTAL = OFF_MBAR + G::STAGES * 8;
    static_assert(OFF_MSK % 16 == 0 && OFF_CMP % 16 == 0, "alignment");
};

__device__ __forceinline__ uint2 lds64(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
// bit scans: the device intrinsics, with host stand-ins so that tests/native can run the field split on the CPU
// (ctz of 0 is the word width, ffs of 0 is 0, as on the device)
__host__ __device__ __forceinline__ uint32_t ctz32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return (uint32_t)__clz((int)__brev(x));
#else
    return x ? (uint32_t)__builtin_ctz(x) : 32u;
#endif
}
__host__ __device__ __forceinline__ uint32_t ctz64(uint64_t x) {
#ifdef __CUDA_ARCH__
    return (uint32_t)__clzll((long long)__brevll(x));
#else
    return x ? (uint32_t)__builtin_ctzll(x) : 64u;
#endif
}
__host__ __device__ __forceinline__ uint32_t l2_ffs32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return (uint32_t)__ffs((int)x);
#else
    return x ? (uint32_t)__builtin_ctz(x) + 1u : 0u;
#endif
}
__host__ __device__ __forceinline__ uint32_t l2_ffs64(uint64_t x) {
#ifdef __CUDA_ARCH__
    return (uint32_t)__ffsll((long long)x);
#else
    return x ? (uint32_t)__builtin_ctzll(x) + 1u : 0u;
#endif
}
__host__ __device__ __forceinline__ uint32_t bfind32(uint32_t x) {  // index of the highest set bit
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
    return r;
#else
    return x ? 31u - (uint32_t)__builtin_clz(x) : 0xffffffffu;
#endif
}

// A decimal of <= 4 digits as its digit bytes, most significant on top (d3 << 24 | d2 << 16 | d1 << 8 | d0):
// monotone in the value, so intervals of values are intervals of codes and no multiply is needed.
__host__ __device__ __forceinline__ uint32_t lean2_code(uint32_t v) {  // v <= 9999
    return ((v / 1000u) << 24) | (((v / 100u) % 10u) << 16) | (((v / 10u) % 10u) << 8) | (v % 10u);
}

// the complement of the modular interval {lo, width} (v - lo <= width, unsigned): again one
__host__ __device__ __forceinline__ void lean2_negate(uint32_t& lo, uint32_t& width) {
    if (width == 0xffffffffu) {  // always -> never
        lo = 0xffffffffu;
        width = 0u;
    } else {
        lo = lo + width + 1u;
        width = 0xfffffffeu - width;
    }
}

// `mant * A <op> LB` over mant in [0, 2^24) as a modular interval: pass = (mant - lo) <= width (unsigned).
// "never" is {0xffffffff, 0}, "always" {0, 0xffffffff}; != is the complement of the == interval.
// Also the same test over digit codes (lean2_code) of values <= 9999: {clo, cwidth}.
__host__ __device__ inline void lean2_interval(const LeanLeaf& L, int fd, uint32_t& lo, uint32_t& width, uint32_t& clo, uint32_t& cwidth) {
    uint32_t negate;
    const unsigned long long A = L.A[fd] ? L.A[fd] : 1u;
    const long long LB = L.LB[fd];
    const unsigned long long top = 0xfffffffeull;
    lo = 0u;
    width = 0xffffffffu;
    negate = 0u;
    if (L.lop == 0) {  // mant * A > LB
        if (LB >= 0) {
            const unsigned long long q = (unsigned long long)LB / A;
            if (q >= top) {
                negate = 1u;
            } else {
                lo = (uint32_t)q + 1u;
                width = 0xffffffffu - lo;
            }
        }
    } else if (L.lop == 1) {  // mant * A < LB
        if (LB <= 0) {
            negate = 1u;
        } else {
            const unsigned long long q = ((unsigned long long)LB + A - 1ull) / A;  // mant <= q - 1
            if (q - 1ull < top) width = (uint32_t)(q - 1ull);
        }
    } else {  // == / !=
        const bool hit = LB >= 0 && (unsigned long long)LB % A == 0ull && (unsigned long long)LB / A <= top;
        if (hit) {
            lo = (uint32_t)((unsigned long long)LB / A);
            width = 0u;
            negate = L.lop == 3 ? 1u : 0u;
        } else {
            negate = L.lop == 2 ? 1u : 0u;
        }
    }
    if (negate && L.lop != 3) {
        // (only whole-range intervals are negated here: "never")
        lo = 0xffffffffu;
        width = 0u;
        negate = 0u;
    }
    if (L.lop == 3 && !negate) {  // != without an exact hit: always true
        lo = 0u;
        width = 0xffffffffu;
    }
    // codes: [lo, lo + width] cut to 0..9999
    clo = 0xffffffffu;
    cwidth = 0u;
    if (lo <= 9999u) {
        const uint32_t hi_v = (width >= 9999u - lo) ? 9999u : lo + width;
        clo = lean2_code(lo);
        cwidth = lean2_code(hi_v) - clo;
        if (lo == 0u && hi_v == 9999u) cwidth = 0xffffffffu;
    }
    if (L.lop == 3 && negate) {  // != with an exact hit: the complement of {E, 0}
        lean2_negate(lo, width);
        lean2_negate(clo, cwidth);
    }
}

// the same for a field with a leading '-': `-(mant / 10^fd) <op> literal` over mant, i.e. mant * A <op mirrored> -LB
__host__ __device__ inline void lean2_interval_neg(const LeanLeaf& L, int fd, uint32_t& lo, uint32_t& width) {
    LeanLeaf M = L;
    M.LB[fd] = -L.LB[fd];
    M.lop = L.lop == 0 ? 1 : L.lop == 1 ? 0 : L.lop;
    uint32_t clo, cwidth;
    lean2_interval(M, fd, lo, width, clo, cwidth);
}

// a + c as an IMAD on the device (add_fma, cqg_lean.cuh), plainly on the host
__host__ __device__ __forceinline__ uint32_t l2_add(uint32_t a, uint32_t one, uint32_t c) {
#ifdef __CUDA_ARCH__
    return add_fma(a, one, c);
#else
    return a * one + c;
#endif
}
// the three integer intrinsics of the decode, with host stand-ins so that tests/native can run the same arithmetic
__host__ __device__ __forceinline__ uint32_t l2_prmt(uint32_t a, uint32_t b, uint32_t s) {
#ifdef __CUDA_ARCH__
    return __byte_perm(a, b, s);
#else
    const unsigned long long pool = ((unsigned long long)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((pool >> (8 * ((s >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
    return r;
#endif
}
__host__ __device__ __forceinline__ uint32_t l2_umulhi(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((unsigned long long)a * b) >> 32);
#endif
}
__host__ __device__ __forceinline__ uint32_t l2_dp4a(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
    return __dp4a(a, b, c);
#else
    for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 0xffu) * ((b >> (8 * i)) & 0xffu);
    return c;
#endif
}

// phase 1 of the scalar lean kernel on one 16-byte chunk: bit i of `t16` = byte i is special (a control byte, so every
// terminator, or the quote '"'; kLeanSpecialXor), bit i of `d16` = byte i is the delimiter (patD = it, in every byte). Bytes with
// bit 7 set (UTF-8) are neither. 3 instructions per word and class, then the 0x80 flags become mask bits
// through four dot products per class (weights 1,2,4,8 | 16..128 leave mask << 7).
// `kx`: kLeanSpecialXor, handed in so that the kernel can pin it in a register ((v ^ kx) & 0x7f.. is then ONE LOP3).
__host__ __device__ __forceinline__ void lean2_masks16(uint32_t vx, uint32_t vy, uint32_t vz, uint32_t vw, uint32_t patD, uint32_t one,
                                                       uint32_t& t16, uint32_t& d16, uint32_t kx = kLeanSpecialXor) {
    const uint32_t a0 = ~(l2_add((vx ^ kx) & 0x7f7f7f7fu, one, 0x5f5f5f5fu) | vx) & 0x80808080u;
    const uint32_t a1 = ~(l2_add((vy ^ kx) & 0x7f7f7f7fu, one, 0x5f5f5f5fu) | vy) & 0x80808080u;
    const uint32_t a2 = ~(l2_add((vz ^ kx) & 0x7f7f7f7fu, one, 0x5f5f5f5fu) | vz) & 0x80808080u;
    const uint32_t a3 = ~(l2_add((vw ^ kx) & 0x7f7f7f7fu, one, 0x5f5f5f5fu) | vw) & 0x80808080u;
    const uint32_t d0 = ~(l2_add((vx ^ patD) & 0x7f7f7f7fu, one, 0x7f7f7f7fu) | vx) & 0x80808080u;
    const uint32_t d1 = ~(l2_add((vy ^ patD) & 0x7f7f7f7fu, one, 0x7f7f7f7fu) | vy) & 0x80808080u;
    const uint32_t d2 = ~(l2_add((vz ^ patD) & 0x7f7f7f7fu, one, 0x7f7f7f7fu) | vz) & 0x80808080u;
    const uint32_t d3 = ~(l2_add((vw ^ patD) & 0x7f7f7f7fu, one, 0x7f7f7f7fu) | vw) & 0x80808080u;
    uint32_t ra = l2_dp4a(a2, 0x08040201u, 0u);
    ra = l2_dp4a(a3, 0x80402010u, ra) * 256u;
    ra = l2_dp4a(a0, 0x08040201u, ra);
    ra = l2_dp4a(a1, 0x80402010u, ra);
    uint32_t rd = l2_dp4a(d2, 0x08040201u, 0u);
    rd = l2_dp4a(d3, 0x80402010u, rd) * 256u;
    rd = l2_dp4a(d0, 0x08040201u, rd);
    rd = l2_dp4a(d1, 0x80402010u, rd);
    t16 = ra >> 7;
    d16 = rd >> 7;
}

// unsigned decimal of 1..4 bytes given as the 4 bytes ENDING at the field's end (`w`: last character on top,
// whatever precedes the field below): mant / 10^(fd16/16). false: not one (sign, exponent, text, two dots, a lone '.').
__host__ __device__ __forceinline__ bool lean2_dec4_word(uint32_t w, uint32_t len, uint32_t& mant, uint32_t& fd16) {
    uint32_t t = w ^ 0x30303030u;
    t &= ~(0x00ffffffu >> (8u * len - 8u));                        // what precedes the field reads as leading zeros
    const uint32_t x = ((t ^ 0x1e1e1e1eu) & 0x7f7f7f7fu) + 0x7f7f7f7fu;
    const uint32_t dotf = ~(x | t) & 0x80808080u;                  // 0x80 where the byte is '.'
    fd16 = 0u;
    if (dotf) {
        if ((dotf & (dotf - 1u)) || len == 1u) return false;
        // fd = 3 - (byte index of the dot), as fd * 0x11: nibble 0 picks the low selector byte, nibble 1 the high one
        const uint32_t r = l2_umulhi(dotf, 0x66442200u);
        fd16 = r & 0x30u;
        const uint32_t sel = l2_prmt(0x14040404u, 0x32323121u, (r & 0x33u) | 0x40u);
        t = l2_prmt(t, 0u, sel);                                   // drop the dot, shift the integer digits down
    }
    if (((t + 0x76767676u) | t) & 0x80808080u) return false;      // a byte that is not a digit
    mant = l2_dp4a(t, 0x010a6400u, (t & 0xffu) * 1000u);
    return true;
}
// the field ENDS at shared address `fe` (exclusive)
__device__ __forceinline__ bool lean2_dec4(uint32_t fe, uint32_t len, uint32_t& mant, uint32_t& fd16) {
    const uint32_t a = fe & ~3u;
    const uint32_t w0 = lds32(a - 4u), w1 = lds32(a);
    return lean2_dec4_word(__funnelshift_r(w0, w1, fe << 3), len, mant, fd16);  // bytes [fe-4, fe)
}
// lean2_dec4_word with the digit code (lean2_code) instead of the value (the byte permute that drops the '.' also
// reverses the bytes). Returns 0 when the field is such a decimal, anything else when not. `kdot` = 0x1e1e1e1e
// ('.' ^ '0' in every byte), handed in so that the kernel can pin it in a register.
__host__ __device__ __forceinline__ uint32_t lean2_dec4c_word(uint32_t w, uint32_t len, uint32_t kdot, uint32_t& code, uint32_t& fd16) {
    uint32_t t = w ^ 0x30303030u;
    t &= ~(0x00ffffffu >> (8u * len - 8u));
    const uint32_t x = ((t ^ kdot) & 0x7f7f7f7fu) + 0x7f7f7f7fu;
    const uint32_t dotf = ~(x | t) & 0x80808080u;
    fd16 = 0u;
    uint32_t sel = 0x0123u, bad = 0u;
    if (dotf) {
        bad = (dotf & (dotf - 1u)) | (len == 1u ? 1u : 0u);
        const uint32_t r = l2_umulhi(dotf, 0x66442200u);           // fd * 0x11 (see lean2_dec4_word)
        fd16 = r & 0x30u;
        sel = l2_prmt(0x23231312u, 0x41404040u, (r & 0x33u) | 0x40u);
    }
    t = l2_prmt(t, 0u, sel);
    bad |= ((t + 0x76767676u) | t) & 0x80808080u;                 // a byte that is not a digit
    code = t;
    return bad;
}
__device__ __forceinline__ uint32_t lean2_dec4c(uint32_t fe, uint32_t len, uint32_t kdot, uint32_t& code, uint32_t& fd16) {
    const uint32_t a = fe & ~3u;
    const uint32_t w0 = lds32(a - 4u), w1 = lds32(a);
    return lean2_dec4c_word(__funnelshift_r(w0, w1, fe << 3), len, kdot, code, fd16);
}
// first terminator bit at or after bit `q` of the tile buffer; `limit` when there is none below it
__device__ __noinline__ uint32_t lean2_next_term(uint32_t s_msk, uint32_t q, uint32_t limit) {
    uint32_t w = q >> 5;
    uint32_t t = lds32(s_msk + 8u * w) & (0xffffffffu << (q & 31u));
    while (t == 0u && (w + 1u) * 32u < limit) {
        w++;
        t = lds32(s_msk + 8u * w);
    }
    const uint32_t e = w * 32u + ctz32(t);
    return (t != 0u && e < limit) ? e : limit;
}

// field positions of one row out of its stop bits (delimiters below the terminator, and the terminator):
// start (relative to the row) and length of wanted field K; a field the row does not have reads as empty
template <typename W>
struct Lean2Stops {
    W st;
    uint32_t sp;
    bool missing;
    __host__ __device__ __forceinline__ void field(int gap, uint32_t& off, uint32_t& len) {
        if (gap > 0) {
            CQG_SPEC_UNROLL
            for (int i = 1; i < gap; i++) st &= st - 1;
            missing = missing || st == 0;
            sp = sizeof(W) == 8 ? l2_ffs64((uint64_t)st) : l2_ffs32((uint32_t)st);
            st &= st - 1;
        }
        missing = missing || st == 0;
        const uint32_t ep = sizeof(W) == 8 ? ctz64((uint64_t)st) : ctz32((uint32_t)st);
        off = sp;
        len = missing ? 0u : ep - sp;
    }
};


// One wanted field (column GAP0, or gap0 when GAP0 < 0) of a row that ends inside its 32-bit window: tw / dw = the
// terminator / delimiter bits from the row's first byte on. et = index of the terminator, sp = start of the field,
// flen = its length. The stops are the row's delimiters, its terminator, and every bit above it as a sentinel, so
// that a field the row does not have comes out with length 0 (terminator on bit 31: 32 - sp, 0 or far too long).
template <int GAP0>
__host__ __device__ __forceinline__ void lean2_oneleaf_field(uint32_t tw, uint32_t dw, int gap0, uint32_t& et, uint32_t& sp, uint32_t& flen) {
    const uint32_t below = tw ^ (tw - 1u);  // up to and including the terminator
    et = bfind32(below);
    uint32_t st = dw | tw | ~below;
    if (GAP0 > 0) {
#pragma unroll
        for (int i = 1; i < GAP0; i++) st &= st - 1u;
        sp = l2_ffs32(st);
        st &= st - 1u;
    } else if (GAP0 < 0 && gap0 > 0) {
#pragma unroll 1
        for (int i = 1; i < gap0; i++) st &= st - 1u;
        sp = l2_ffs32(st);
        st &= st - 1u;
    }
    flen = ctz32(st) - sp;
}

// A row that does not end inside the 32-bit window (ONELEAF): 32..63 bytes on 64-bit masks, longer ones are
// found and handed over. Returns et | sp << 16 | flen << 24 | dirty << 31 (flen 0: hand the row over).
__device__ __noinline__ uint32_t lean2_wide_row(uint32_t s_msk, uint32_t pos, int gap, uint32_t limit) {
    const uint32_t ma = s_msk + ((pos >> 2) & ~7u);
    const uint2 m0 = lds64(ma), m1 = lds64(ma + 8u), m2 = lds64(ma + 16u);
    const uint32_t tw2 = __funnelshift_r(m1.x, m2.x, pos);
    if (tw2 == 0u) {
        const uint32_t e = lean2_next_term(s_msk, pos + 64u, limit);
        return (e - pos) | (e >= limit ? 0x80000000u : 0u);
    }
    const uint64_t tw64 = (uint64_t)tw2 << 32;
    const uint64_t dw64 = ((uint64_t)__funnelshift_r(m1.y, m2.y, pos) << 32) | __funnelshift_r(m0.y, m1.y, pos);
    const uint64_t below = tw64 ^ (tw64 - 1ull);
    const uint32_t et = 32u + bfind32((uint32_t)(below >> 32));
    uint64_t st = (dw64 | tw64) & below;
    uint32_t sp = 0;
    if (gap > 0) {
        for (int i = 1; i < gap; i++) st &= st - 1ull;
        sp = (uint32_t)__ffsll((long long)st);
        st &= st - 1ull;
    }
    const uint32_t flen = st ? ctz64(st) - sp : 0u;
    return et | (sp << 16) | ((flen > 63u ? 0u : flen) << 24);
}

// unsigned decimal of 5..7 bytes (rare next to the 4-byte route: kept out of line)
// returns mant (< 10^7) | fd16 << 24 | ok << 31
__device__ __noinline__ uint32_t lean2_dec7(uint32_t fa, uint32_t len) {
    uint32_t mant, fd;
    bool hd;
    const bool ok = lean_decimal(fa, len, mant, fd, hd);
    return ok ? (0x80000000u | (fd << 28) | mant) : 0u;
}
// A decimal with a leading sign, 2..7 bytes in all ("-5", "+1.25", "-123.4"): shorter than the 8 bytes at which the
// reference tries a date first (infer_type, src/csv_reader.c:137), typed and valued as strtoll / strtod do
// (src/csv_reader.c:160-193, :195-240). Returns mant (< 10^6) | negative << 27 | fd << 28 | ok << 31. Out of line: a
// field reaches it only after the unsigned decode said no.
__device__ __noinline__ uint32_t lean2_signed(uint32_t fa, uint32_t len) {
    if (len - 2u > 5u) return 0u;
    const uint32_t c0 = lds8(fa);
    if (c0 != '-' && c0 != '+') return 0u;
    uint32_t mant, fd;
    bool hd;
    if (!lean_decimal(fa + 1u, len - 1u, mant, fd, hd)) return 0u;
    return 0x80000000u | (fd << 28) | (c0 == '-' ? 0x08000000u : 0u) | mant;
}
// The sign travels in bit 3 of fd16 (fd16 = 16 * fraction digits | 8 * negative): a leaf's interval table holds, 8 bytes
// behind the interval for `mant / 10^fd <op> literal`, the one for `-(mant / 10^fd) <op> literal` (lean2_interval_neg), so
// a comparison costs a signed field nothing extra; sums mask the bit off and negate.
#define CQG_L2_SIGNED(FA, LEN, DEC, MANT, FD16)                               \
    if (!(DEC)) {                                                             \
        const uint32_t rs_ = lean2_signed(FA, LEN);                           \
        if (rs_ >> 31) {                                                      \
            DEC = true;                                                       \
            MANT = rs_ & 0x00ffffffu;                                         \
            FD16 = (rs_ >> 24) & 0x38u;                                       \
        }                                                                     \
    }
// value * 1000 of a decoded field as a two's-complement 64-bit integer
__device__ __forceinline__ unsigned long long lean2_times_1000(uint32_t mant, uint32_t fd16) {
    const uint32_t f = fd16 & 0x30u;
    const unsigned long long v = (unsigned long long)mant * (f == 0u ? 1000u : f == 16u ? 100u : f == 32u ? 10u : 1u);
    return (fd16 & 8u) ? 0ull - v : v;
}
#define CQG_L2_DEC7(FA, LEN, DEC, MANT, FD16)        \
    {                                                \
        const uint32_t r7 = lean2_dec7(FA, LEN);     \
        DEC = (r7 >> 31) != 0u;                      \
        MANT = r7 & 0x00ffffffu;                     \
        FD16 = (r7 >> 24) & 0x30u;                   \
    }

// decimal of the field in slot SL (1..7 bytes; DEC stays false otherwise). A kernel compiled for one query knows
// its slots, so a column read by a leaf AND an aggregate (`WHERE age > 25 ... AVG(age)`) is decoded once per row.
#ifdef CQG_JIT
#define CQG_L2_DECODE_STATE uint32_t dcache_mant[4] = {0u, 0u, 0u, 0u}, dcache_fd[4] = {0u, 0u, 0u, 0u}, dcache_state = 0u;
#define CQG_L2_DECODE(SL, RB, O, L, DEC, MANT, FD16)                                        \
    {                                                                                       \
        const uint32_t st_ = (dcache_state >> (2 * (SL))) & 3u;                             \
        if (st_ != 0u) {                                                                    \
            DEC = st_ == 1u;                                                                \
            MANT = dcache_mant[SL];                                                         \
            FD16 = dcache_fd[SL];                                                           \
        } else {                                                                            \
            if ((L) - 1u < 4u) {                                                            \
                DEC = lean2_dec4((RB) + (O) + (L), L, MANT, FD16);                          \
            } else if ((L) - 1u < 7u) {                                                     \
                CQG_L2_DEC7((RB) + (O), L, DEC, MANT, FD16)                                 \
            }                                                                               \
            CQG_L2_SIGNED((RB) + (O), L, DEC, MANT, FD16)                                   \
            dcache_mant[SL] = MANT;                                                         \
            dcache_fd[SL] = FD16;                                                           \
            dcache_state |= (DEC ? 1u : 2u) << (2 * (SL));                                  \
        }                                                                                   \
    }
#else
#define CQG_L2_DECODE_STATE
#define CQG_L2_DECODE(SL, RB, O, L, DEC, MANT, FD16)                                        \
    {                                                                                       \
        if ((L) - 1u < 4u) {                                                                \
            DEC = lean2_dec4((RB) + (O) + (L), L, MANT, FD16);                              \
        } else if ((L) - 1u < 7u) {                                                         \
            CQG_L2_DEC7((RB) + (O), L, DEC, MANT, FD16)                                     \
        }                                                                                   \
        CQG_L2_SIGNED((RB) + (O), L, DEC, MANT, FD16)                                       \
    }
#endif

// ---- tiles at the edges of the file (DevPlan::edge_in_kernel) ----
// The first tile has no PRE bytes in front of it and the last ones run past the end of the file: the lean kernels used to
// leave both to the general kernel (one more launch and one more host round trip per query: 0.06 ms, 2 % of a 10 GB
// COUNT and 6 % of a 1.25 GB slice on 8 GPUs). With the flag set (scans of >= 64 tiles) the part of the tile that exists
// is loaded - 16-byte multiples by the bulk copy, up to 15 bytes more one by one - and everything else reads as '\n',
// which is how csv_load sees the edges too (a file starts at a line start; its last line needs no terminator,
// src/csv_reader.c:404-427).
struct LeanEdge {
    uint32_t lo_b, hi_b, load16;  // buffer bytes [lo_b, hi_b) exist in the file; [lo_b, lo_b + load16) come by bulk copy
};
template <class G>
__device__ __forceinline__ LeanEdge lean_edge_span(long long g0, uint64_t size) {
    LeanEdge e;
    e.lo_b = g0 < 0 ? (uint32_t)(-g0) : 0u;
    const long long avail = (long long)size - g0;
    e.hi_b = avail >= (long long)G::BUF ? (uint32_t)G::BUF : (avail > (long long)e.lo_b ? (uint32_t)avail : e.lo_b);
    e.load16 = (e.hi_b - e.lo_b) & ~15u;
    return e;
}
template <class G>
__device__ __forceinline__ void lean_edge_fill(uint8_t* buf, const uint8_t* data, long long g0, const LeanEdge& e, int tid) {
    for (uint32_t k = (uint32_t)tid; k < (uint32_t)G::BUF; k += (uint32_t)G::THREADS) {
        if (k < e.lo_b || k >= e.hi_b) buf[k] = (uint8_t)'\n';
        else if (k >= e.lo_b + e.load16) buf[k] = data[g0 + (long long)k];
    }
}

// GAP0: the wanted column index of ONELEAF plans when it is below 8 (the delimiter skips unroll), else -1
// CRLF (DevPlan::crlf: the head of the file holds a CR): every line ends in the two bytes CR LF. Both are `T` bytes; a row
// is then clean when its first terminator is the '\r' and the byte behind it the '\n', and the next row starts two bytes
// on. A line that ends in anything else makes the tile dirty as before (a bare '\n' included: the general kernel splits
// such a file exactly as csv_load does, src/csv_reader.c:404-427).
template <class G, int MINB, bool ONELEAF, int GAP0, bool CRLF = false>
__global__ void __launch_bounds__(G::THREADS, MINB) lean2_kernel(const __grid_constant__ DevPlan P) {
    extern __shared__ __align__(128) uint8_t smem[];
    static_assert(G::STAGES == 1 && G::TILE == G::THREADS * 128, "one stage, 128 bytes per thread");
    using LL = Lean2Layout<G>;
    uint32_t sbase;
    asm volatile("mov.u32 %0, %1;" : "=r"(sbase) : "r"(smem_u32(smem)));  // once: not to be rematerialised per row
    const uint32_t s_buf = sbase + G::OFF_BUF, s_msk = sbase + LL::OFF_MSK, s_cmp = sbase + LL::OFF_CMP;
    uint64_t* mbar = (uint64_t*)(smem + LL::OFF_MBAR);
    unsigned* s_handed = (unsigned*)(smem + LL::OFF_MISC);
    const int tid = threadIdx.x, lane = tid & 31;

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_handed[0] = 0u;
        s_handed[1] = 0u;
    }
    for (int w = G::BUF / 32 + tid; w < G::MASKW; w += G::THREADS) {  // behind the buffer: all terminators
        sts32(s_msk + 8 * w, 0xffffffffu);
        sts32(s_msk + 8 * w + 4, 0u);
    }
    if (tid < P.l_nleaf * 4 && P.l_leaf[tid >> 2].kind == 0) {
        uint32_t lo, width, clo, cwidth;
        lean2_interval(P.l_leaf[tid >> 2], tid & 3, lo, width, clo, cwidth);
        if (!ONELEAF) lean2_interval_neg(P.l_leaf[tid >> 2], tid & 3, clo, cwidth);  // (ONELEAF: digit-code intervals there)
        sts64(s_cmp + 16 * tid, lo, width);
        sts64(s_cmp + 16 * tid + 8, clo, cwidth);
    }
    __syncthreads();

    // running totals of clean tiles
    uint32_t rows = 0, count = 0;
    uint64_t first = ~0ull;
    bool have_first = false;
    long long s3[4];
    uint32_t sn[4];
#pragma unroll
    for (int a = 0; a < 4; a++) {
        s3[a] = 0;
        sn[a] = 0;
    }
    const uint64_t size = P.size;
    const uint32_t patD = (uint32_t)P.delim * 0x01010101u;
    const uint32_t one = (uint32_t)P.simple;  // == 1, but not to the compiler: a + c as IMAD (FMA pipe), see add_fma
    uint32_t kxor;
    asm volatile("mov.u32 %0, %1;" : "=r"(kxor) : "n"(kLeanSpecialXor));  // opaque: stays in a register
    const int nwant = CQG_SPEC(NWANT, P.nwantL);
    const int gap0 = GAP0 >= 0 ? GAP0 : CQG_SPEC(GAP0, P.gap[0]), gap1 = CQG_SPEC(GAP1, P.gap[1]), gap2 = CQG_SPEC(GAP2, P.gap[2]),
              gap3 = CQG_SPEC(GAP3, P.gap[3]);
    const int nprog = CQG_SPEC(NPROG, P.l_nprog);
    const int nagg = CQG_SPEC(NAGG, P.l_nagg);
    uint32_t summask = 0;
    int aslot[4];
#pragma unroll
    for (int a = 0; a < 4; a++) {
        aslot[a] = 0;
        if (!ONELEAF && a < nagg) {
            summask |= 1u << a;
            aslot[a] = CQG_SPEC_AT(ASLOT, a, P.aggs[P.l_agg[a]].slot);
        }
    }

    const int my_tiles = (P.n_tiles > (int)blockIdx.x) ? (P.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    for (int it = 0; it < my_tiles; it++) {
        const long long tile = (long long)P.first_tile + blockIdx.x + (long long)it * gridDim.x;
        const long long g0 = tile * (long long)G::TILE - G::PRE;
        const bool at_edge = g0 < 0 || g0 + G::BUF > (long long)size;
        const bool edge = at_edge && !P.edge_in_kernel;  // handed over, not loaded
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
            // this CTA's next tile: on its way into L2 while this one is worked on (one stage of shared memory only)
            const long long gn = g0 + (long long)gridDim.x * G::TILE;
            if (it + 1 < my_tiles && gn + G::BUF <= (long long)size)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(P.data + gn), "r"((uint32_t)G::BUF) : "memory");
            s_handed[it & 1] = 0u;  // last read two tiles ago, at least one barrier back
        }
        mbar_wait(&mbar[0], (uint32_t)it & 1u);
        if (edge) {
            if (tid == 0) {
                unsigned long long k = atomicAdd(P.def_tile_count, 1ull);
                P.def_tiles[k] = (int32_t)tile;
            }
            __syncthreads();
            continue;
        }
        if (at_edge) {
            lean_edge_fill<G>(smem + G::OFF_BUF, P.data, g0, es, tid);
            __syncthreads();
        }

        // ---- phase 1: T ("< 0x23") and D (delimiter) masks, 16 bytes per thread and step ----
        {
            const uint32_t ca = s_buf + 16u * tid;
            const uint32_t ma = s_msk + (((uint32_t)tid >> 1) << 3) + (((uint32_t)tid & 1u) << 1);
            auto chunk = [&](uint32_t ca, uint32_t ma) {
                const uint4 v = lds128(ca);
                uint32_t ra, rd;
                lean2_masks16(v.x, v.y, v.z, v.w, patD, one, ra, rd, kxor);
                sts16(ma, ra);
                sts16(ma + 4u, rd);
            };
            constexpr int kFull = G::CHUNKS / G::THREADS;  // steps every thread takes
#pragma unroll
            for (int k = 0; k < kFull; k++) chunk(ca + 16u * G::THREADS * k, ma + 4u * G::THREADS * k);
            if (tid < G::CHUNKS - kFull * G::THREADS) chunk(ca + 16u * G::THREADS * kFull, ma + 4u * G::THREADS * kFull);
        }
        __syncthreads();

        // ---- phase 2: every thread walks the rows that START in its own 128 bytes ----
        uint32_t lo = (uint32_t)G::PRE + 128u * (uint32_t)tid, hi = lo + 128u;
        {
            const long long olo_l = (long long)P.own_lo - g0, ohi_l = (long long)P.own_hi - g0;
            if (olo_l > (long long)G::PRE || ohi_l < (long long)(G::PRE + G::TILE)) {  // first / last tile of a shard
                const uint32_t olo = (uint32_t)(olo_l < G::PRE ? G::PRE : (olo_l > G::PRE + G::TILE ? G::PRE + G::TILE : olo_l));
                const uint32_t ohi = (uint32_t)(ohi_l < G::PRE ? G::PRE : (ohi_l > G::PRE + G::TILE ? G::PRE + G::TILE : ohi_l));
                lo = lo > olo ? lo : olo;
                hi = hi < ohi ? hi : ohi;
            }
        }
        uint32_t tcnt = 0, trows = 0, tfirst = 0xffffffffu, seen = 0, dirty = 0, nh = 0, hpos = 0;
        long long ts3[4];
        uint32_t tsn[4];
#pragma unroll
        for (int a = 0; a < 4; a++) {
            ts3[a] = 0;
            tsn[a] = 0;
        }
        if (lo < hi) {
            uint32_t e0;
            {
                const uint32_t q = lo - 1u;
                const uint32_t qa = s_msk + ((q >> 2) & ~7u);
                const uint32_t fw = __funnelshift_r(lds32(qa), lds32(qa + 8u), q);
                e0 = fw != 0u ? q + ctz32(fw) : lean2_next_term(s_msk, q + 32u, (uint32_t)G::BUF);
            }
            uint32_t pos = e0 + 1u;
            if (CRLF && pos < hi && lds8(s_buf + e0) == 0x0du) pos = e0 + 2u;  // (e0 is the CR of the pair, or its LF)
            if (pos < hi) {
                if (CRLF) dirty |= (lds8(s_buf + pos - 2u) != 0x0du || lds8(s_buf + pos - 1u) != 0x0au) ? 1u : 0u;
                else dirty |= lds8(s_buf + e0) != 0x0au ? 1u : 0u;
                uint32_t ma = s_msk + ((pos >> 2) & ~7u);
                uint2 m0 = lds64(ma), m1 = lds64(ma + 8u);
                if constexpr (ONELEAF) {
                    // ---- COUNT(*) WHERE column <op> literal: the row loop written out for this shape alone ----
                    uint32_t iters = 0, nlacc = 0;
                    uint32_t kdot;  // '.' ^ '0' in every byte, pinned in a register (a second LOP3 immediate otherwise)
                    asm volatile("mov.u32 %0, 0x1e1e1e1e;" : "=r"(kdot));
                    do {
                        const uint32_t tw = __funnelshift_r(m0.x, m1.x, pos);
                        const uint32_t dw = __funnelshift_r(m0.y, m1.y, pos);
                        uint32_t et, sp = 0, flen;
                        if (tw != 0u) {
                            seen |= tw;
                            lean2_oneleaf_field<GAP0>(tw, dw, gap0, et, sp, flen);
                        } else {
                            const uint32_t r = lean2_wide_row(s_msk, pos, gap0, (uint32_t)G::BUF);
                            et = r & 0xffffu;
                            sp = (r >> 16) & 0xffu;
                            flen = (r >> 24) & 0x7fu;
                            dirty |= r >> 31;
                        }
                        const uint32_t rbase = s_buf + pos;
                        const uint32_t npos = pos + et + (CRLF ? 2u : 1u);
                        const uint32_t nma = s_msk + ((npos >> 2) & ~7u);
                        uint2 n0, n1;
                        uint32_t lastb;
                        n0 = lds64(nma);  // (asked for before the decode, used after it)
                        n1 = lds64(nma + 8u);
                        lastb = lds8(rbase + et);
                        if (CRLF) lastb = (lastb ^ 0x07u) | (lds8(rbase + et + 1u) ^ 0x0au) << 8;  // 0x0a when the line ends in CR LF
                        uint32_t val = 0, fd16 = 0, tab = 8u, bad = 1u;
                        if (flen - 1u < 4u) {
                            bad = lean2_dec4c(rbase + sp + flen, flen, kdot, val, fd16);
                        } else if (flen - 1u < 7u) {
                            bool dec;
                            CQG_L2_DEC7(rbase + sp, flen, dec, val, fd16)
                            bad = dec ? 0u : 1u;
                            tab = 0u;
                        }
                        const uint2 iv = lds64(s_cmp + fd16 + tab);
                        // count the row, keep the first one, or note it for the general kernel (the last two: pos < 2^16)
                        asm("{\n\t"
                            ".reg .pred good, take;\n\t"
                            ".reg .u32 d;\n\t"
                            "setp.eq.u32 good, %4, 0;\n\t"
                            "sub.u32 d, %5, %6;\n\t"
                            "setp.le.and.u32 take, d, %7, good;\n\t"
                            "@take add.u32 %0, %0, 1;\n\t"
                            "@take min.u32 %1, %1, %8;\n\t"
                            "@!good mad.lo.u32 %2, %2, 65536, %8;\n\t"
                            "@!good add.u32 %3, %3, 1;\n\t"
                            "}"
                            : "+r"(tcnt), "+r"(tfirst), "+r"(hpos), "+r"(nh)
                            : "r"(bad), "r"(val), "r"(iv.x), "r"(iv.y), "r"(pos));
                        nlacc |= lastb ^ 0x0au;  // the row must end in '\n'
                        iters++;
                        pos = npos;
                        ma = nma;
                        m0 = n0;
                        m1 = n1;
                    } while (pos < hi);
                    trows = iters - nh;
                    dirty |= nlacc != 0u ? 1u : 0u;
                } else
                do {
                    const uint32_t tw = __funnelshift_r(m0.x, m1.x, pos);
                    const uint32_t dw = __funnelshift_r(m0.y, m1.y, pos);
                    bool ok = true, pass = true;
                    uint32_t et;
                    uint32_t off0 = 0, off1 = 0, off2 = 0, off3 = 0, len0 = 0, len1 = 0, len2 = 0, len3 = 0;
                    if (tw != 0u) {
                        // the row ends inside a 32-bit window
                        seen |= tw;
                        const uint32_t below = tw ^ (tw - 1u);  // up to and including the terminator
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
                            const uint32_t d2 = lds32(ma + 20u);
                            const uint64_t tw64 = (uint64_t)tw2 << 32;
                            const uint64_t dw64 = ((uint64_t)__funnelshift_r(m1.y, d2, pos) << 32) | dw;
                            const uint64_t below = tw64 ^ (tw64 - 1ull);
                            et = 32u + bfind32((uint32_t)(below >> 32));
                            Lean2Stops<uint64_t> S{(dw64 | tw64) & below, 0u, false};
                            S.field(gap0, off0, len0);
                            if (nwant > 1) S.field(gap1, off1, len1);
                            if (nwant > 2) S.field(gap2, off2, len2);
                            if (nwant > 3) S.field(gap3, off3, len3);
                        } else {
                            // 64 bytes or more: find the end, hand the row over
                            const uint32_t e = lean2_next_term(s_msk, pos + 64u, (uint32_t)G::BUF);
                            if (e >= (uint32_t)G::BUF) dirty = 1u;
                            et = e - pos;
                            ok = false;
                        }
                    }
                    const uint32_t rbase = s_buf + pos;
                    // the next row's mask words and this row's last byte: asked for now, used after the decode
                    const uint32_t npos = pos + et + (CRLF ? 2u : 1u);
                    const uint32_t nma = s_msk + ((npos >> 2) & ~7u);
                    uint2 n0, n1;
                    uint32_t lastb;
                    n0 = lds64(nma);
                    n1 = lds64(nma + 8u);
                    lastb = lds8(rbase + et);
                    if (CRLF) lastb = (lastb ^ 0x07u) | (lds8(rbase + et + 1u) ^ 0x0au) << 8;  // 0x0a when the line ends in CR LF
                    unsigned long long add0 = 0, add1 = 0, add2 = 0, add3 = 0;
                    uint32_t addmask = 0;
                    CQG_L2_DECODE_STATE
                    if (ok) {
#define CQG_L2_SLOT(SL, O, L)                                                    \
    const uint32_t O = SL == 0 ? off0 : SL == 1 ? off1 : SL == 2 ? off2 : off3; \
    const uint32_t L = SL == 0 ? len0 : SL == 1 ? len1 : SL == 2 ? len2 : len3;
                        if (nprog) {
                            uint32_t bs = 0;
                            CQG_SPEC_UNROLL
                            for (int pc = 0; pc < nprog; pc++) {
                                const int c = CQG_SPEC_AT(PROG, pc, P.l_prog[pc]);
                                if (c >= 0) {
                                    const int sl = CQG_SPEC_AT(LEAFSLOT, c, P.l_leaf[c].slot), kind = CQG_SPEC_AT(LEAFKIND, c, P.l_leaf[c].kind);
                                    CQG_L2_SLOT(sl, o, l)
                                    bool bv = false;
                                    if (kind == 0) {
                                        uint32_t mant = 0, fd16 = 0;
                                        bool dec = false;
                                        CQG_L2_DECODE(sl, rbase, o, l, dec, mant, fd16)
                                        if (dec) {
                                            const uint2 iv = lds64(s_cmp + 64u * (uint32_t)c + fd16);
                                            bv = mant - iv.x <= iv.y;
                                        } else {
                                            ok = false;  // NULL, text, date, signed or long number: general kernel
                                        }
                                    } else {
                                        // text equality (see cqg_lean.cuh: a number or date against text is handed over)
                                        uint32_t tag;
                                        uint64_t w0, w1;
                                        if (l == 0u) {
                                            bv = kind == 2;
                                        } else if (l > 16u) {
                                            const uint32_t c0 = lds8(rbase + o);
                                            const bool ns = (c0 - 48u) <= 9u || c0 == '+' || c0 == '-' || c0 == '.';
                                            if (ns || c0 == ' ' || lds8(rbase + o + l - 1u) == ' ') ok = false;  // (trimmed by the reference)
                                            bv = kind == 2;
                                        } else if (lean_key_part(rbase + o, l, tag, w0, w1) && (tag == KT_STR || tag == KT_NULL)) {
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
                        if (ok && pass && summask) {
#define CQG_L2_AGG(A, ADD)                                                                               \
    if (summask & (1u << A)) {                                                                           \
        const int sl = aslot[A];                                                                         \
        CQG_L2_SLOT(sl, o, l)                                                                            \
        uint32_t mant = 0, fd16 = 0;                                                                     \
        bool dec = false;                                                                                \
        CQG_L2_DECODE(sl, rbase, o, l, dec, mant, fd16) \
        if (dec) {                                                                                       \
            ADD = lean2_times_1000(mant, fd16);                                                          \
            addmask |= 1u << A;                                                                          \
        } else if (l != 0u) {                                                                            \
            ok = false; /* a value this kernel does not decode (NULL is simply not summed) */           \
        }                                                                                                \
    }
                            CQG_L2_AGG(0, add0)
                            CQG_L2_AGG(1, add1)
                            CQG_L2_AGG(2, add2)
                            CQG_L2_AGG(3, add3)
#undef CQG_L2_AGG
                        }
#undef CQG_L2_SLOT
                    }
                    if (ok) {
                        trows++;
                        if (pass) {
                            tcnt++;
                            tfirst = tfirst < pos ? tfirst : pos;
                            if (!ONELEAF) {
                                if (addmask & 1u) {
                                    ts3[0] += (long long)add0;
                                    tsn[0]++;
                                }
                                if (addmask & 2u) {
                                    ts3[1] += (long long)add1;
                                    tsn[1]++;
                                }
                                if (addmask & 4u) {
                                    ts3[2] += (long long)add2;
                                    tsn[2]++;
                                }
                                if (addmask & 8u) {
                                    ts3[3] += (long long)add3;
                                    tsn[3]++;
                                }
                            }
                        }
                    } else {
                        hpos = hpos * 65536u + pos;  // the last two rows handed over (pos < 2^16)
                        nh++;
                    }
                    // the row must end in '\n' (anything else below 0x23 is not this kernel's business)
                    dirty |= lastb != 0x0au ? 1u : 0u;
                    pos = npos;
                    ma = nma;
                    m0 = n0;
                    m1 = n1;
                } while (pos < hi);
                dirty |= seen & 1u;  // a row start that is a terminator: an empty line
            }
        }
        if (nh > 2u) dirty = 1u;
        if (nh != 0u && dirty == 0u) atomicAdd(&s_handed[it & 1], nh);
        int bad = __syncthreads_or((int)dirty);
        // (this barrier also ends every read of the tile and its masks)
        if (s_handed[it & 1] > 32u) bad = 1;
        if (!bad) {
            rows += trows;
            count += tcnt;
            if (tcnt != 0u && !have_first) {
                first = P.global_base + (uint64_t)(g0 + (long long)tfirst);
                have_first = true;
            }
#pragma unroll
            for (int a = 0; a < 4; a++) {
                s3[a] += ts3[a];
                sn[a] += tsn[a];
            }
            if (nh != 0u) {
                const unsigned long long k = atomicAdd(P.def_row_count, (unsigned long long)nh);
                if (k < P.def_row_cap) P.def_rows[k] = (uint64_t)(g0 + (long long)(hpos & 0xffffu));
                if (nh > 1u && k + 1ull < P.def_row_cap) P.def_rows[k + 1ull] = (uint64_t)(g0 + (long long)(hpos >> 16));
            }
        } else if (tid == 0) {
            unsigned long long k = atomicAdd(P.def_tile_count, 1ull);
            P.def_tiles[k] = (int32_t)tile;
        }
    }

    // ---- epilogue: fold the registers into the single group `_all_` ----
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        rows += __shfl_xor_sync(0xffffffffu, rows, d);
        count += __shfl_xor_sync(0xffffffffu, count, d);
        const uint64_t of = __shfl_xor_sync(0xffffffffu, first, d);
        first = of < first ? of : first;
#pragma unroll
        for (int a = 0; a < 4; a++) {
            if (!ONELEAF) {
                s3[a] += __shfl_xor_sync(0xffffffffu, s3[a], d);
                sn[a] += __shfl_xor_sync(0xffffffffu, sn[a], d);
            }
        }
    }
    if (lane == 0 && rows) atomicAdd(P.rows_scanned, (unsigned long long)rows);
    if (lane == 0 && count) {
        unsigned err = 0;
        const uint64_t h = key_hash_final(0x243F6A8885A308D3ull);
        uint8_t* ge = global_entry_for(P, h, 0u, nullptr, err);
        if (ge) {
            atomicAdd((unsigned long long*)(ge + kOffCount), (unsigned long long)count);
            amin64((uint64_t*)(ge + kOffFirst), first << 16);
#pragma unroll
            for (int a = 0; a < 4; a++) {
                if (!ONELEAF && a < nagg && sn[a]) {
                    atomicAdd((unsigned long long*)(ge + P.aggs[P.l_agg[a]].off + 16), (unsigned long long)sn[a]);
                    atomicAdd((unsigned long long*)(ge + P.aggs[P.l_agg[a]].off + 24), (unsigned long long)s3[a]);
                }
            }
        }
        if (err) atomicOr(P.errflags, err);
    }
}

}  // namespace cqg
