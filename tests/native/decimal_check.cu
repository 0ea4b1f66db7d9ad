// Host-side exhaustive check of the 4-byte decimal decode of the lean kernels (cqg_lean2.cuh: lean2_dec4_word,
// lean2_dec4c_word): every string of 1..4 characters over an alphabet of digits and everything that may stand
// next to them, with every possible garbage in the bytes in front of the field. Expected: what parse_value
// (src/csv_reader.c:195-240) makes of an unsigned decimal without exponent — value = digits / 10^(digits behind the
// dot) — and "not covered" for everything else (signs, exponents, text, two dots, a lone dot).
// Built and run by tests/test_interval_logic.py (host code, no GPU).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "cqg_lean2.cuh"

using namespace cqg;

int main() {
    const char alphabet[] = "0123456789.+-eE ,xn\"\n\xc3";
    const int na = (int)sizeof(alphabet) - 1;
    const unsigned char fillers[] = {',', '\n', '9', '.', 0xff, 0x00};
    long long checked = 0;
    for (int len = 1; len <= 4; len++) {
        long long combos = 1;
        for (int i = 0; i < len; i++) combos *= na;
        for (long long c = 0; c < combos; c++) {
            unsigned char f[4];
            long long v = c;
            for (int i = 0; i < len; i++) {
                f[i] = (unsigned char)alphabet[v % na];
                v /= na;
            }
            // expected
            int dots = 0, digits = 0, fd = 0;
            unsigned mant = 0;
            bool other = false;
            for (int i = 0; i < len; i++) {
                if (f[i] >= '0' && f[i] <= '9') {
                    digits++;
                    mant = mant * 10 + (f[i] - '0');
                    if (dots) fd++;
                } else if (f[i] == '.') {
                    dots++;
                } else {
                    other = true;
                }
            }
            const bool want = !other && dots <= 1 && digits >= 1;
            for (unsigned char fill : fillers) {
                uint32_t w = 0;  // the 4 bytes ending at the field's end: last character in the top byte
                for (int i = 0; i < 4; i++) {
                    const int k = len - 4 + i;  // index into the field of byte i of the window
                    const unsigned char b = k >= 0 ? f[k] : fill;
                    w |= (uint32_t)b << (8 * i);
                }
                uint32_t m1 = 0, fd1 = 0, code = 0, fd2 = 0;
                const bool ok1 = lean2_dec4_word(w, (uint32_t)len, m1, fd1);
                const bool ok2 = lean2_dec4c_word(w, (uint32_t)len, 0x1e1e1e1eu, code, fd2) == 0u;
                if (ok1 != want || ok2 != want || (want && (m1 != mant || fd1 != 16u * fd || fd2 != 16u * fd || code != lean2_code(mant)))) {
                    printf("MISMATCH '%.*s' fill %02x: want %d mant %u fd %d; value route %d %u %u; code route %d %08x %u\n", len, (const char*)f,
                           fill, (int)want, mant, fd, (int)ok1, m1, fd1 / 16, (int)ok2, code, fd2 / 16);
                    return 1;
                }
                checked++;
            }
        }
    }
    printf("ok %lld fields\n", checked);
    return 0;
}
