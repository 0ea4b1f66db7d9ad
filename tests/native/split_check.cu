// Host-side check of the lean kernels' field split (cqg_lean2.cuh: Lean2Stops): wanted fields of random rows of
// 0..62 bytes, taken off the row's delimiter / terminator bitmasks with `x &= x - 1` skips, against a plain
// byte-by-byte split (parse_line without quotes, src/csv_reader.c:278-338). Rows with fewer fields than asked
// for must come out with length 0 for the missing ones. Built and run by tests/test_interval_logic.py (host code).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cqg_lean2.cuh"

using namespace cqg;

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return (uint32_t)(rng_state >> 24);
}

template <typename W>
static long long run(int max_len, int rows) {
    long long checked = 0;
    for (int r = 0; r < rows; r++) {
        // a row: nf fields of random (often zero) lengths, joined by ',', ended by '\n'
        const int nf = 1 + (int)(rnd() % 7);
        std::vector<int> flen(nf);
        int total = nf - 1;
        for (int f = 0; f < nf; f++) {
            flen[f] = (rnd() % 4 == 0) ? 0 : (int)(rnd() % 9);
            total += flen[f];
        }
        if (total + 1 > max_len) continue;
        // masks as phase 1 leaves them: bit i = byte i of the row; bytes behind the terminator belong to later rows
        W dw = 0, tw = 0;
        int pos = 0;
        std::vector<int> start(nf);
        for (int f = 0; f < nf; f++) {
            start[f] = pos;
            pos += flen[f];
            if (f + 1 < nf) dw |= (W)1 << pos++;
        }
        tw |= (W)1 << pos;
        for (int i = pos + 1; i < (int)(8 * sizeof(W)); i++) {  // the next rows: arbitrary delimiters and terminators
            const uint32_t x = rnd() % 8;
            if (x == 0) dw |= (W)1 << i;
            if (x == 1) tw |= (W)1 << i;
        }
        // up to four wanted columns, ascending, possibly beyond the row's last field
        int want[4], nwant = 1 + (int)(rnd() % 4), c = (int)(rnd() % 3);
        for (int k = 0; k < nwant; k++) {
            want[k] = c;
            c += 1 + (int)(rnd() % 3);
        }
        const W below = tw ^ (tw - 1);
        Lean2Stops<W> S{(W)((dw | tw) & below), 0u, false};
        for (int k = 0; k < nwant; k++) {
            uint32_t off = 0, len = 0;
            S.field(k == 0 ? want[0] : want[k] - want[k - 1], off, len);
            const bool have = want[k] < nf;
            const uint32_t want_len = have ? (uint32_t)flen[want[k]] : 0u;
            if (len != want_len || (have && want_len && off != (uint32_t)start[want[k]])) {
                printf("MISMATCH row of %d fields, wanted column %d: off %u len %u, expected off %d len %u\n", nf, want[k], off, len,
                       have ? start[want[k]] : -1, want_len);
                exit(1);
            }
            checked++;
        }
    }
    return checked;
}

// the COUNT(*)-WHERE kernel's own single-field version (stops with sentinel bits, compile-time or run-time column)
template <int GAP0>
static long long run_oneleaf(int rows, int gap_runtime) {
    long long checked = 0;
    const int col = GAP0 >= 0 ? GAP0 : gap_runtime;
    for (int r = 0; r < rows; r++) {
        const int nf = 1 + (int)(rnd() % 11);
        std::vector<int> flen(nf), start(nf);
        int total = nf - 1;
        for (int f = 0; f < nf; f++) {
            flen[f] = (rnd() % 4 == 0) ? 0 : (int)(rnd() % 6);
            total += flen[f];
        }
        if (total + 1 > 32) continue;
        uint32_t dw = 0, tw = 0;
        int pos = 0;
        for (int f = 0; f < nf; f++) {
            start[f] = pos;
            pos += flen[f];
            if (f + 1 < nf) dw |= 1u << pos++;
        }
        tw |= 1u << pos;
        for (int i = pos + 1; i < 32; i++) {
            const uint32_t x = rnd() % 8;
            if (x == 0) dw |= 1u << i;
            if (x == 1) tw |= 1u << i;
        }
        uint32_t et = 0, sp = 0, fl = 0;
        lean2_oneleaf_field<GAP0>(tw, dw, gap_runtime, et, sp, fl);
        const bool have = col < nf;
        // a missing field must not look like a decimal of 1..7 bytes (the kernel hands such rows over)
        const bool bad = et != (uint32_t)pos || (have ? (fl != (uint32_t)flen[col] || (fl && sp != (uint32_t)start[col])) : (fl - 1u < 7u));
        if (bad) {
            printf("MISMATCH oneleaf column %d of %d fields: et %u sp %u len %u, expected et %d start %d len %d\n", col, nf, et, sp, fl, pos,
                   have ? start[col] : -1, have ? flen[col] : -1);
            exit(1);
        }
        checked++;
    }
    return checked;
}

int main() {
    long long o = 0;
    o += run_oneleaf<0>(100000, 0) + run_oneleaf<1>(100000, 1) + run_oneleaf<2>(100000, 2) + run_oneleaf<3>(100000, 3);
    o += run_oneleaf<5>(100000, 5) + run_oneleaf<7>(100000, 7);
    for (int g = 0; g < 12; g++) o += run_oneleaf<-1>(50000, g);
    printf("ok %lld single-field rows; ", o);
    const long long a = run<uint32_t>(32, 400000);
    const long long b = run<uint64_t>(64, 400000);
    printf("%lld fields on 32-bit masks, %lld on 64-bit masks\n", a, b);
    return 0;
}
