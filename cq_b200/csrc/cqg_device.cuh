// cqg_device.cuh — device-side value model of the cq hot path: typed decode of a CSV field,
// value_compare, LIKE, arithmetic and the predicate interpreter. Hand-written for sm_100a;
// no libc on this side, so every libc call of the reference (strtoll / strtod / sscanf /
// isspace / tolower / strcmp) is restated here and cited.
#pragma once
#include "cqg_rtc.h"
#ifndef __CUDACC_RTC__
#include <cstdint>
#include <cuda_runtime.h>
#endif

#include "cq_gpu.h"

namespace cqg {

// ---- error flags raised by kernels (host turns them into CQG_ERR_UNSUPPORTED) ----
enum : unsigned {
    KERR_NUMERIC_RANGE = 1u,   // decimal with >19 significant digits or >19 fraction digits
    KERR_TABLE_FULL = 2u,      // group table over capacity: host retries with a bigger one
    KERR_SEP_OVERFLOW = 4u,    // separator list of a tile overflowed shared memory
    KERR_KEY_RANGE = 8u,       // DOUBLE group key outside the %.6f integer range handled
    KERR_MINMAX_TIE = 16u,     // INTEGER/DOUBLE tie in MIN/MAX: order-dependent result type
    KERR_STACK = 32u,          // predicate stack overflow
    KERR_SEL_OVERFLOW = 64u,   // selection buffer full: host retries with a bigger one
    KERR_JOIN_MIXED = 128u,    // join key column mixes comparison classes (Q7 cross-type equal)
    KERR_BIGINT = 256u         // |INTEGER| > 2^53 where the reference compares as double
};

constexpr int T_NULL = CQG_TYPE_NULL, T_INT = CQG_TYPE_INTEGER, T_DBL = CQG_TYPE_DOUBLE,
              T_STR = CQG_TYPE_STRING, T_DATE = CQG_TYPE_DATE, T_BOOL = 100;

struct DVal {
    int32_t type;
    uint32_t len;  // STRING: trimmed length
    union {
        long long i;  // INTEGER; DATE packed (y<<16 | m<<8 | d); BOOL
        double d;
        const uint8_t* s;  // STRING: trimmed view (generic address: smem tile, HBM, const pool)
    };
};

__host__ __device__ __forceinline__ bool is_space(uint32_t c) { return c == 32u || (c - 9u) <= 4u; }  // C-locale isspace
__host__ __device__ __forceinline__ bool is_digit(uint32_t c) { return (c - 48u) <= 9u; }
__host__ __device__ __forceinline__ uint32_t to_lower(uint32_t c) { return (c - 65u) <= 25u ? c + 32u : c; }

// ------------------------------------------------------------------------------------------
// dates: src/date_utils.c:8-100
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ bool valid_date(long long y, long long m, long long d) {
    if (y < 1000 || y > 9999) return false;
    if (m < 1 || m > 12) return false;
    if (d < 1) return false;
    int dim = (m == 2) ? ((((y % 4 == 0) && (y % 100 != 0)) || (y % 400 == 0)) ? 29 : 28)
                       : ((m == 4 || m == 6 || m == 9 || m == 11) ? 30 : 31);
    return d <= dim;
}

// one sscanf("%d") conversion on [i, n): skip isspace, optional sign, >=1 digit, at most
// `width` characters (sign included, as glibc counts it). Returns false on matching failure.
__host__ __device__ __forceinline__ bool scan_int(const uint8_t* p, uint32_t& i, uint32_t n, uint32_t width,
                                                   long long& out) {
    while (i < n && is_space(p[i])) i++;
    bool neg = false;
    uint32_t used = 0;
    if (i < n && (p[i] == '+' || p[i] == '-') && used < width) {
        neg = p[i] == '-';
        i++;
        used++;
    }
    long long v = 0;
    uint32_t nd = 0;
    while (i < n && used < width && is_digit(p[i])) {
        v = v * 10 + (long long)(p[i] - 48u);  // <= 10 digits: no overflow
        i++;
        used++;
        nd++;
    }
    if (nd == 0) return false;
    long long sv = neg ? -v : v;
    out = (long long)(int)sv;  // stored through an int* (date_utils.c:29)
    return true;
}

// parse_date on the field [p, p+len), 8 <= len <= 10, after the isspace trim of
// src/csv_reader.c:143-149. Returns packed date or -1.
__host__ __device__ inline long long try_date(const uint8_t* p, uint32_t len) {
    uint32_t a = 0, n = len;
    while (a < n && is_space(p[a])) a++;
    while (n > a && is_space(p[n - 1])) n--;
    // a C string ends at a NUL byte
    for (uint32_t k = a; k < n; k++)
        if (p[k] == 0) {
            n = k;
            break;
        }
    if (a >= n) return -1;
    uint32_t c0 = p[a];
    if (!(is_digit(c0) || c0 == '+' || c0 == '-')) return -1;  // every format starts with %d
    long long x = 0, y = 0, z = 0;
    // ISO "%d-%d-%d" (date_utils.c:34)
    {
        uint32_t i = a;
        if (scan_int(p, i, n, 0xffffffffu, x) && i < n && p[i] == '-') {
            i++;
            if (scan_int(p, i, n, 0xffffffffu, y) && i < n && p[i] == '-') {
                i++;
                if (scan_int(p, i, n, 0xffffffffu, z) && valid_date(x, y, z)) return (x << 16) | (y << 8) | z;
            }
        }
    }
    // US "%d/%d/%d" as m/d/y (:46) then EU as d/m/y (:58)
    {
        uint32_t i = a;
        if (scan_int(p, i, n, 0xffffffffu, x) && i < n && p[i] == '/') {
            i++;
            if (scan_int(p, i, n, 0xffffffffu, y) && i < n && p[i] == '/') {
                i++;
                if (scan_int(p, i, n, 0xffffffffu, z)) {
                    if (valid_date(z, x, y)) return (z << 16) | (x << 8) | y;
                    if (valid_date(z, y, x)) return (z << 16) | (y << 8) | x;
                }
            }
        }
    }
    // COMPACT "%8d" (:70-76); C division truncates toward zero
    {
        uint32_t i = a;
        if (scan_int(p, i, n, 8, x)) {
            long long d = x % 100;
            long long t = x / 100;
            long long m = t % 100;
            long long yy = t / 100;
            if (valid_date(yy, m, d)) return (yy << 16) | (m << 8) | d;
        }
    }
    return -1;
}

// ------------------------------------------------------------------------------------------
// numbers: strtoll / strtod on the grammar infer_type admits (src/csv_reader.c:158-193)
// ------------------------------------------------------------------------------------------
__device__ __constant__ double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                             1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

__device__ __forceinline__ unsigned long long pow10_u64(int k) {
    unsigned long long r = 1;
    for (int j = 0; j < k; j++) r *= 10ull;
    return r;
}

// correctly rounded (nearest-even) mant / 10^fd for fd <= 19, any mant: 128-bit long division,
// then rounding of the 128-bit quotient plus sticky bit. This is what glibc strtod returns.
__device__ __noinline__ double div_pow10_exact(unsigned long long mant, int fd) {
    if (mant == 0) return 0.0;
    unsigned long long D = pow10_u64(fd);
    unsigned long long hi = mant / D, r = mant % D, lo = 0;
    // lo = floor(r * 2^64 / D) by shift-subtract (r < D < 2^64)
    for (int b = 0; b < 64; b++) {
        bool carry = (r >> 63) != 0;
        r <<= 1;
        lo <<= 1;
        if (carry || r >= D) {
            r -= D;
            lo |= 1;
        }
    }
    bool sticky = r != 0;
    // value = (hi*2^64 + lo [+sticky]) * 2^-64 ; normalise to 64 significant bits
    int e;  // value = m64 * 2^e with bit 63 of m64 set
    unsigned long long m64;
    if (hi) {
        int lz = __clzll((long long)hi);
        m64 = lz ? ((hi << lz) | (lo >> (64 - lz))) : hi;
        unsigned long long rest = lz ? (lo << lz) : lo;
        sticky = sticky || rest != 0;
        e = -lz;
    } else {
        int lz = __clzll((long long)lo);
        m64 = lo << lz;
        e = -64 - lz;
    }
    // round 64 -> 53 bits
    unsigned long long q = m64 >> 11, rem = m64 & 0x7ffull;
    bool up = rem > 0x400ull || (rem == 0x400ull && (sticky || (q & 1ull)));
    if (up) {
        q++;
        if (q >> 53) {
            q >>= 1;
            e++;
        }
    }
    // q in [2^52, 2^53): double = q * 2^(e+11); all results here are normal numbers
    int ex = e + 11 + 52 + 1023;
    unsigned long long bits = ((unsigned long long)ex << 52) | (q & 0xfffffffffffffull);
    return __longlong_as_double((long long)bits);
}

// ------------------------------------------------------------------------------------------
// strtod for decimal strings of any length (src/csv_reader.c:210 calls glibc strtod, which is
// correctly rounded): big-integer long division, used when the 64-bit paths above do not apply.
// text = [digits][.digits] (sign and blanks already consumed), n bytes.
// ------------------------------------------------------------------------------------------
constexpr int kBigLimbs = 52;     // 1664 bits
constexpr int kBigMaxDigits = 400;

__device__ __forceinline__ int big_bitlen(const uint32_t* a) {
    for (int i = kBigLimbs - 1; i >= 0; i--)
        if (a[i]) return 32 * i + (32 - __clz((int)a[i]));
    return 0;
}
__device__ __forceinline__ void big_mul10_add(uint32_t* a, uint32_t d) {
    unsigned long long carry = d;
    for (int i = 0; i < kBigLimbs; i++) {
        unsigned long long t = (unsigned long long)a[i] * 10ull + carry;
        a[i] = (uint32_t)t;
        carry = t >> 32;
    }
}
__device__ __forceinline__ void big_shl(uint32_t* a, int s) {  // s >= 0, result must fit
    int w = s >> 5, b = s & 31;
    for (int i = kBigLimbs - 1; i >= 0; i--) {
        uint32_t lo = (i - w >= 0) ? a[i - w] : 0u;
        uint32_t lo2 = (i - w - 1 >= 0) ? a[i - w - 1] : 0u;
        a[i] = b ? ((lo << b) | (lo2 >> (32 - b))) : lo;
    }
}
__device__ __forceinline__ void big_shr1(uint32_t* a) {
    for (int i = 0; i < kBigLimbs; i++) a[i] = (a[i] >> 1) | (i + 1 < kBigLimbs ? (a[i + 1] << 31) : 0u);
}
__device__ __forceinline__ bool big_ge(const uint32_t* a, const uint32_t* b) {
    for (int i = kBigLimbs - 1; i >= 0; i--)
        if (a[i] != b[i]) return a[i] > b[i];
    return true;
}
__device__ __forceinline__ void big_sub(uint32_t* a, const uint32_t* b) {  // a -= b, a >= b
    unsigned long long borrow = 0;
    for (int i = 0; i < kBigLimbs; i++) {
        unsigned long long t = (unsigned long long)a[i] - b[i] - borrow;
        a[i] = (uint32_t)t;
        borrow = (t >> 32) & 1ull;
    }
}

__device__ __noinline__ double strtod_big(const uint8_t* p, uint32_t n, unsigned& errflags) {
    uint32_t A[kBigLimbs], B[kBigLimbs];
    for (int i = 0; i < kBigLimbs; i++) {
        A[i] = 0;
        B[i] = 0;
    }
    B[0] = 1;
    int ndig = 0, fd = 0;
    bool dot = false;
    // trailing zeros of the fraction do not change the value: leave them out
    uint32_t end = n;
    bool has_dot = false;
    for (uint32_t k = 0; k < n; k++) has_dot = has_dot || p[k] == '.';
    if (has_dot)
        while (end > 0 && p[end - 1] == '0') end--;
    for (uint32_t k = 0; k < end; k++) {
        uint32_t c = p[k];
        if (c == '.') {
            dot = true;
            continue;
        }
        uint32_t d = c - 48u;
        if (ndig == 0 && d == 0) {  // leading zero: no digit of M yet
            if (dot) fd++;
            continue;
        }
        if (ndig >= kBigMaxDigits || fd >= kBigMaxDigits) {
            errflags |= KERR_NUMERIC_RANGE;
            return 0.0;
        }
        big_mul10_add(A, d);
        ndig++;
        if (dot) fd++;
    }
    if (ndig == 0) return 0.0;
    for (int k = 0; k < fd; k++) big_mul10_add(B, 0);
    int la = big_bitlen(A), lb = big_bitlen(B);
    int s = 64 + lb - la;  // 2^63 <= floor(A * 2^s / B) < 2^65
    if (s >= 0) big_shl(A, s);
    else big_shl(B, -s);
    big_shl(B, 64);
    unsigned long long q = 0;
    bool qtop = false;  // bit 64 of the quotient
    for (int i = 64; i >= 0; i--) {
        if (big_ge(A, B)) {
            big_sub(A, B);
            if (i == 64) qtop = true;
            else q |= 1ull << i;
        }
        big_shr1(B);
    }
    bool sticky = false;
    for (int i = 0; i < kBigLimbs; i++) sticky = sticky || A[i] != 0;
    // value = (qtop:q) * 2^-s ; round to 53 bits, nearest even
    int drop = qtop ? 12 : 11;
    unsigned long long m, rem, half = 1ull << (drop - 1);
    if (qtop) {
        m = (q >> 12) | (1ull << 52);
        rem = q & 0xfffull;
    } else {
        m = q >> 11;
        rem = q & 0x7ffull;
    }
    if (rem > half || (rem == half && (sticky || (m & 1ull)))) {
        m++;
        if (m >> 53) {
            m >>= 1;
            drop++;
        }
    }
    int e2 = drop - s;  // value = m * 2^e2, m in [2^52, 2^53)
    int ex = e2 + 52 + 1023;
    if (ex >= 2047) return __longlong_as_double(0x7ff0000000000000ll);  // HUGE_VAL
    if (ex <= 0) {  // subnormal result: different rounding position, not handled
        errflags |= KERR_NUMERIC_RANGE;
        return 0.0;
    }
    return __longlong_as_double((long long)(((unsigned long long)ex << 52) | (m & 0xfffffffffffffull)));
}

// Decode one field exactly as parse_value does (src/csv_reader.c:195-240).
// `errflags` collects KERR_* for inputs outside the exact range handled on the device.
__device__ __noinline__ DVal decode_field(const uint8_t* p, uint32_t len, unsigned& errflags) {
    DVal v;
    v.type = T_NULL;
    v.len = 0;
    v.i = 0;
    if (len == 0) return v;
    if (len >= 8 && len <= 10) {  // :137
        long long dt = try_date(p, len);
        if (dt >= 0) {
            v.type = T_DATE;
            v.i = dt;
            return v;
        }
    }
    uint32_t i = 0;
    while (i < len && is_space(p[i])) i++;
    bool neg = false;
    uint32_t j = i;
    if (j < len && (p[j] == '+' || p[j] == '-')) {
        neg = p[j] == '-';
        j++;
    }
    bool number = j < len;
    bool has_dot = false, has_digit = false;
    unsigned long long mant = 0;
    int nsig = 0, fd = 0;
    bool int_overflow = false;  // an integer-part digit beyond the 19 kept: magnitude changed
    bool inexact = false;       // a non-zero fraction digit beyond the 19 kept
    uint32_t k0 = j;            // end of the number text
    if (number) {
        uint32_t k = j;
        while (k < len && !is_space(p[k])) {
            uint32_t c = p[k];
            if (is_digit(c)) {
                has_digit = true;
                uint32_t dg = c - 48u;
                if (nsig < 19) {
                    if (mant != 0 || dg != 0) {
                        mant = mant * 10ull + dg;
                        nsig++;
                    }
                    if (has_dot) fd++;
                } else if (!has_dot) {
                    int_overflow = true;
                } else if (dg) {
                    inexact = true;  // trailing fraction zeros beyond 19 digits change nothing
                }
            } else if (c == '.' && !has_dot) {
                has_dot = true;
            } else {
                number = false;
                break;
            }
            k++;
        }
        k0 = k;
        if (number) {
            while (k < len && is_space(p[k])) k++;
            number = has_digit && k == len;
        }
    }
    if (number && !has_dot) {
        // strtoll(str, NULL, 10): saturates (:207)
        v.type = T_INT;
        bool sat = int_overflow;
        if (!sat) {
            if (!neg && mant > 0x7fffffffffffffffull) sat = true;
            if (neg && mant > 0x8000000000000000ull) sat = true;
        }
        if (sat) v.i = neg ? (long long)0x8000000000000000ull : 0x7fffffffffffffffll;
        else v.i = neg ? (long long)(0ull - mant) : (long long)mant;
        return v;
    }
    if (number) {
        // strtod (:210), correctly rounded
        v.type = T_DBL;
        double r;
        if (!(int_overflow || inexact) && mant < (1ull << 53) && fd <= 22) {
            r = (double)(long long)mant / kPow10[fd];  // one correctly rounded IEEE division
        } else if (!(int_overflow || inexact) && fd <= 19) {
            r = div_pow10_exact(mant, fd);
        } else {
            uint32_t e = k0;  // end of the number text (before trailing blanks)
            r = strtod_big(p + j, e - j, errflags);
        }
        v.d = neg ? -r : r;
        return v;
    }
    // STRING: cq_strndup + trim_whitespace (:234-235) as a view
    v.type = T_STR;
    uint32_t a = 0, n = len;
    for (uint32_t k = 0; k < n; k++)
        if (p[k] == 0) {  // a NUL ends the reference's C string
            n = k;
            break;
        }
    while (a < n && is_space(p[a])) a++;
    while (n > a + 1 && is_space(p[n - 1])) n--;
    v.s = p + a;
    v.len = n - a;
    return v;
}

// parse_value for a field of a CLEAN tile: the tile holds no byte below 0x23 other than '\n'
// (no blanks, no NUL, no quote), so trimming and NUL handling of src/csv_reader.c:143-149,
// :234-235 are no-ops and the numeric grammar of infer_type (:158-193) is [sign](digit|one dot)+.
// Anything this routine does not settle in a few instructions goes to decode_field.
__device__ __forceinline__ DVal decode_field_clean(const uint8_t* p, uint32_t len, unsigned& errflags) {
    DVal v;
    v.type = T_NULL;
    v.len = 0;
    v.i = 0;
    if (len == 0) return v;
    uint32_t c0 = p[0];
    bool numeric_start = is_digit(c0) || c0 == '+' || c0 == '-' || c0 == '.';
    if (!numeric_start) {
        // cannot be a date (every sscanf format starts with %d) nor a number
        v.type = T_STR;
        v.s = p;
        v.len = len;
        return v;
    }
    if (len > 7) return decode_field(p, len, errflags);  // dates (8..10 bytes) and long numbers
    bool neg = c0 == '-';
    uint32_t mant = 0, fd = 0;
    bool dot = false, digit = false, ok = true;
    for (uint32_t k = (c0 == '+' || c0 == '-') ? 1u : 0u; k < len; k++) {  // len <= 7: mant < 10^7
        uint32_t c = p[k];
        uint32_t d = c - 48u;
        if (d <= 9u) {
            mant = mant * 10u + d;
            digit = true;
            fd += dot ? 1u : 0u;
        } else {
            ok = ok && c == '.' && !dot;
            dot = true;
        }
    }
    if (!(ok && digit)) {  // "-", "+", ".", "1-2", "12ab": STRING (no blanks to trim in a clean tile)
        v.type = T_STR;
        v.s = p;
        v.len = len;
        return v;
    }
    if (!dot) {
        v.type = T_INT;
        v.i = neg ? -(long long)mant : (long long)mant;
        return v;
    }
    v.type = T_DBL;
    double r = (double)mant / kPow10[fd];  // mant < 10^7: one correctly rounded division = strtod
    v.d = neg ? -r : r;
    return v;
}

// ------------------------------------------------------------------------------------------
// value_compare (src/csv_reader.c:98-130)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int str_cmp(const uint8_t* a, uint32_t al, const uint8_t* b, uint32_t bl) {
    uint32_t n = al < bl ? al : bl;
    for (uint32_t k = 0; k < n; k++) {
        int ca = a[k], cb = b[k];
        if (ca != cb) return ca - cb;
    }
    if (al == bl) return 0;
    return al < bl ? -(int)b[n] : (int)a[n];
}

__device__ __forceinline__ double num_of(const DVal& v) { return v.type == T_INT ? (double)v.i : v.d; }

__device__ inline int val_compare(const DVal& a, const DVal& b) {
    if (a.type == T_NULL && b.type == T_NULL) return 0;
    if (a.type == T_NULL) return -1;
    if (b.type == T_NULL) return 1;
    if (a.type == T_DATE && b.type == T_DATE) return a.i < b.i ? -1 : (a.i > b.i ? 1 : 0);
    bool an = a.type == T_INT || a.type == T_DBL, bn = b.type == T_INT || b.type == T_DBL;
    if (an && bn) {
        double x = num_of(a), y = num_of(b);
        return x < y ? -1 : (x > y ? 1 : 0);
    }
    if (a.type == T_STR && b.type == T_STR) return str_cmp(a.s, a.len, b.s, b.len);
    return 0;
}

// match_pattern (src/evaluator/evaluator_conditions.c:16-59)
__device__ inline bool like_match(const uint8_t* str, uint32_t sl, const uint8_t* pat, uint32_t pl, bool cs) {
    uint32_t s = 0, p = 0, star = 0xffffffffu, ss = 0;
    while (s < sl) {
        uint32_t pc = p < pl ? pat[p] : 0u;
        if (pc == '%') {
            star = p++;
            ss = s;
        } else if (pc == '_') {
            s++;
            p++;
        } else {
            bool m = cs ? (str[s] == pc) : (to_lower(str[s]) == to_lower(pc));
            if (m) {
                s++;
                p++;
            } else if (star != 0xffffffffu) {
                p = star + 1;
                s = ++ss;
            } else {
                return false;
            }
        }
    }
    while (p < pl && pat[p] == '%') p++;
    return p == pl;
}

// `(long long)x` as x86-64 computes it (cvttsd2si: 0x8000000000000000 when out of range)
__device__ __forceinline__ long long d2ll_x86(double x) {
    if (!(x > -9223372036854775808.0 && x < 9223372036854775808.0)) {
        if (x == -9223372036854775808.0) return (long long)0x8000000000000000ull;
        return (long long)0x8000000000000000ull;
    }
    return (long long)x;
}

// BINARY_OP arm of evaluate_expression (evaluator_expressions.c:156-262)
__device__ inline DVal arith(int op, const DVal& l, const DVal& r) {
    DVal res;
    res.type = T_NULL;
    res.len = 0;
    res.i = 0;
    bool lint = l.type == T_INT, rint = r.type == T_INT;
    if (!(lint || l.type == T_DBL) || !(rint || r.type == T_DBL)) return res;
    double lv = num_of(l), rv = num_of(r);
    double out = 0.0;
    long long outi = 0;
    bool is_int = false;
    switch (op) {
        case CQG_OP_ADD: out = lv + rv; break;
        case CQG_OP_SUB: out = lv - rv; break;
        case CQG_OP_MUL: out = lv * rv; break;
        case CQG_OP_DIV:
            if (rv == 0) return res;
            out = lv / rv;
            break;
        case CQG_OP_MOD:
            if (lint && rint) {
                if (r.i == 0) return res;
                outi = (r.i == -1) ? 0 : l.i % r.i;
                is_int = true;
            } else {
                if (rv == 0) return res;
                out = fmod(lv, rv);
            }
            break;
        case CQG_OP_BAND:
            if (!(lint && rint)) return res;
            outi = l.i & r.i;
            is_int = true;
            break;
        case CQG_OP_BOR:
            if (!(lint && rint)) return res;
            outi = l.i | r.i;
            is_int = true;
            break;
        case CQG_OP_BXOR:
            if (!(lint && rint)) return res;
            outi = l.i ^ r.i;
            is_int = true;
            break;
        default: break;
    }
    if (is_int) {
        res.type = T_INT;
        res.i = outi;
    } else if (lint && rint && out == (double)d2ll_x86(out)) {
        res.type = T_INT;
        res.i = d2ll_x86(out);
    } else {
        res.type = T_DBL;
        res.d = out;
    }
    return res;
}

// ------------------------------------------------------------------------------------------
// constants of a predicate, staged in device memory
// ------------------------------------------------------------------------------------------
struct DConst {
    int32_t type;
    uint32_t len;
    long long bits;  // INTEGER / DOUBLE bits / DATE packed / STRING: offset into the string pool
};

struct DPred {
    const cqg_insn_t* code;
    int n_code;
    const DConst* consts;
    const uint8_t* pool;
};

constexpr int kStackMax = 24;

// A row as the predicate / aggregation sees it: for every column the plan references, the
// field's location. `colmap[c]` maps a query column index to a slot in fld[] (or -1).
struct RowView {
    const uint8_t* base;   // left row bytes (smem tile or HBM)
    const uint8_t* rbase;  // right row bytes (HBM) for joined rows
    const uint8_t* lfile;  // pointer p such that (field address - p) == offset in the left file
    const uint8_t* rfile;  // same for the right file
    const uint32_t* foff;  // [nslots] field start (relative to base / rbase)
    const uint32_t* flen;  // [nslots] field length; 0 => NULL
    const int16_t* colslot;  // query column -> slot
    int ncols_total;
    int nleft_slots;  // slots < nleft_slots are relative to base, others to rbase
};

__device__ __forceinline__ DVal row_value(const RowView& rv, int col, unsigned& err) {
    DVal v;
    v.type = T_NULL;
    v.len = 0;
    v.i = 0;
    if (col < 0 || col >= rv.ncols_total) return v;
    int s = rv.colslot[col];
    if (s < 0) return v;
    const uint8_t* b = s < rv.nleft_slots ? rv.base : rv.rbase;
    return decode_field(b + rv.foff[s], rv.flen[s], err);
}

// evaluate_condition (evaluator_conditions.c:62-164) over postfix code
__device__ inline bool eval_pred(const DPred& P, const RowView& rv, unsigned& err) {
    if (P.n_code == 0) return true;
    DVal st[kStackMax];
    int sp = 0;
    for (int pc = 0; pc < P.n_code; pc++) {
        int op = P.code[pc].op, a = P.code[pc].a;
        if (sp >= kStackMax - 1 && (op == CQG_OP_COL || op == CQG_OP_CONST || op == CQG_OP_TRUE || op == CQG_OP_FALSE)) {
            err |= KERR_STACK;
            return false;
        }
        switch (op) {
            case CQG_OP_COL: st[sp++] = row_value(rv, a, err); break;
            case CQG_OP_CONST: {
                const DConst c = P.consts[a];
                DVal v;
                v.type = c.type;
                v.len = c.len;
                v.i = c.bits;
                if (c.type == T_STR) v.s = P.pool + c.bits;
                st[sp++] = v;
                break;
            }
            case CQG_OP_ADD: case CQG_OP_SUB: case CQG_OP_MUL: case CQG_OP_DIV: case CQG_OP_MOD:
            case CQG_OP_BAND: case CQG_OP_BOR: case CQG_OP_BXOR: case CQG_OP_ARITH_NULL: {
                DVal r = st[--sp], l = st[--sp];
                st[sp++] = arith(op, l, r);
                break;
            }
            case CQG_OP_NEG: {
                DVal o = st[sp - 1], r;
                r.type = T_NULL;
                r.len = 0;
                r.i = 0;
                if (o.type == T_INT) {
                    r.type = T_INT;
                    r.i = (long long)(0ull - (unsigned long long)o.i);
                } else if (o.type == T_DBL) {
                    r.type = T_DBL;
                    r.d = -o.d;
                }
                st[sp - 1] = r;
                break;
            }
            case CQG_OP_POS: break;
            case CQG_OP_EQ: case CQG_OP_NE: case CQG_OP_GT: case CQG_OP_LT: case CQG_OP_GE: case CQG_OP_LE: {
                DVal r = st[--sp], l = st[--sp];
                int c = val_compare(l, r);
                bool b = op == CQG_OP_EQ ? c == 0 : op == CQG_OP_NE ? c != 0 : op == CQG_OP_GT ? c > 0
                       : op == CQG_OP_LT ? c < 0 : op == CQG_OP_GE ? c >= 0 : c <= 0;
                st[sp].type = T_BOOL;
                st[sp].len = 0;
                st[sp++].i = b;
                break;
            }
            case CQG_OP_IN: case CQG_OP_NOT_IN: {
                bool found = false;
                const DVal& left = st[sp - a - 1];
                for (int k = 0; k < a; k++)
                    if (val_compare(left, st[sp - a + k]) == 0) {
                        found = true;
                        break;
                    }
                sp -= a + 1;
                st[sp].type = T_BOOL;
                st[sp].len = 0;
                st[sp++].i = (op == CQG_OP_IN) ? found : !found;
                break;
            }
            case CQG_OP_LIKE: case CQG_OP_ILIKE: {
                DVal r = st[--sp], l = st[--sp];
                bool b = (l.type == T_STR && r.type == T_STR) ? like_match(l.s, l.len, r.s, r.len, op == CQG_OP_LIKE) : false;
                st[sp].type = T_BOOL;
                st[sp].len = 0;
                st[sp++].i = b;
                break;
            }
            case CQG_OP_AND: {
                bool r = st[--sp].i != 0, l = st[--sp].i != 0;
                st[sp].type = T_BOOL;
                st[sp++].i = l && r;
                break;
            }
            case CQG_OP_OR: {
                bool r = st[--sp].i != 0, l = st[--sp].i != 0;
                st[sp].type = T_BOOL;
                st[sp++].i = l || r;
                break;
            }
            case CQG_OP_NOT: st[sp - 1].i = !(st[sp - 1].i != 0); break;
            case CQG_OP_TRUE: st[sp].type = T_BOOL; st[sp].len = 0; st[sp++].i = 1; break;
            case CQG_OP_FALSE: st[sp].type = T_BOOL; st[sp].len = 0; st[sp++].i = 0; break;
            case CQG_OP_POP: sp--; break;
            default: return false;
        }
    }
    return sp > 0 && st[sp - 1].i != 0;
}

// ------------------------------------------------------------------------------------------
// parse_line restated for one row (src/csv_reader.c:285-338): the exact, sequential splitter.
// Used for rows the mask-based fast path does not cover (quote characters, whitespace
// delimiters, rows longer than a tile window). Records (offset,len) of the columns listed in
// want[] (ascending column indices) into foff/flen; columns not present get len 0.
// ------------------------------------------------------------------------------------------
__device__ inline void split_row_exact(const uint8_t* b, uint32_t rs, uint32_t re, uint8_t delim, uint8_t quote,
                                       const int16_t* want, int nwant, uint32_t* foff, uint32_t* flen) {
    for (int k = 0; k < nwant; k++) {
        foff[k] = rs;
        flen[k] = 0;
    }
    uint32_t ptr = rs;
    int field = 0, wi = 0;
    while (ptr < re && wi < nwant) {
        while (ptr < re && is_space(b[ptr])) ptr++;  // \n and \r cannot occur inside a row
        if (ptr >= re) break;
        uint32_t fs = ptr, fl = 0;
        if (b[ptr] == quote) {
            ptr++;
            fs = ptr;
            while (ptr < re) {
                if (b[ptr] == quote) {
                    if (ptr + 1 < re && b[ptr + 1] == quote) {
                        ptr += 2;
                        fl += 2;
                    } else {
                        fl = ptr - fs;
                        ptr++;
                        break;
                    }
                } else {
                    ptr++;
                }
            }
            while (ptr < re && b[ptr] != delim) ptr++;
        } else {
            while (ptr < re && b[ptr] != delim) ptr++;
            fl = ptr - fs;
        }
        if (want[wi] == field) {
            foff[wi] = fs;
            flen[wi] = fl;
            wi++;
        }
        field++;
        if (ptr < re && b[ptr] == delim) ptr++;
    }
}

// 64-bit mixers
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ unsigned long long hash_bytes(const uint8_t* p, uint32_t n) {
    unsigned long long h = 0x9E3779B97F4A7C15ull ^ n;
    for (uint32_t k = 0; k < n; k++) h = (h ^ p[k]) * 0x100000001B3ull;
    return mix64(h);
}

}  // namespace cqg
