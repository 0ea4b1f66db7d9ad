// Host-side check of phase 1 of the scalar lean kernel (cqg_lean2.cuh: lean2_masks16): the SWAR byte classes and
// the dot-product movemask against a byte loop. Every byte value in every position (next to every byte value in
// the neighbouring position, so that no carry crosses bytes unnoticed), several delimiters, random chunks.
// Built and run by tests/test_interval_logic.py (host code, no GPU).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "cqg_lean2g.cuh"

using namespace cqg;

static uint64_t rng_state = 0x243F6A8885A308D3ull;
static uint32_t rnd() {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return (uint32_t)(rng_state >> 16);
}

static long long g_false_dirty = 0;  // clean chunks the conservative test flags

static long long check(const unsigned char* b, unsigned char delim) {
    uint32_t w[4];
    memcpy(w, b, 16);
    uint32_t t16 = 0, d16 = 0;
    lean2_masks16(w[0], w[1], w[2], w[3], (uint32_t)delim * 0x01010101u, 1u, t16, d16);
    uint32_t et = 0, ed = 0;
    for (int i = 0; i < 16; i++) {
        if (b[i] < 0x20 || b[i] == 0x22) et |= 1u << i;  // special: controls and the quote, not the blank
        if (b[i] == delim) ed |= 1u << i;
    }
    // the GROUP BY kernel's phase 1: exact '\n' class, delimiter class, and "any other byte below 0x23" (which may
    // err on the safe side only: a clean chunk flagged costs time, a dirty one missed would cost the answer)
    uint32_t n16 = 0, d16g = 0, en = 0;
    bool dirty = false;
    const uint32_t spec = l2g_masks16(w[0], w[1], w[2], w[3], (uint32_t)delim * 0x01010101u, 1u, n16, d16g) & 0x80808080u;
    for (int i = 0; i < 16; i++) {
        if (b[i] == '\n') en |= 1u << i;
        if ((b[i] < 0x20 || b[i] == 0x22) && b[i] != '\n') dirty = true;
    }
    if (n16 != en || d16g != ed || (dirty && spec == 0u)) {
        printf("MISMATCH (GROUP BY phase 1) delimiter %02x bytes", delim);
        for (int i = 0; i < 16; i++) printf(" %02x", b[i]);
        printf(": N %04x (expected %04x) D %04x (expected %04x) dirty %d spec %08x\n", n16, en, d16g, ed, (int)dirty, spec);
        exit(1);
    }
    if (!dirty && spec != 0u) g_false_dirty++;
    // the same with CR as a second line terminator (DevPlan::crlf): '\r' joins the terminator class and leaves "dirty"
    {
        uint32_t n16c = 0, d16c = 0, enc = 0;
        bool dirtyc = false;
        const uint32_t specc = l2g_masks16(w[0], w[1], w[2], w[3], (uint32_t)delim * 0x01010101u, 1u, n16c, d16c, true) & 0x80808080u;
        for (int i = 0; i < 16; i++) {
            if (b[i] == '\n' || b[i] == '\r') enc |= 1u << i;
            if ((b[i] < 0x20 || b[i] == 0x22) && b[i] != '\n' && b[i] != '\r') dirtyc = true;
        }
        if (n16c != enc || d16c != ed || (dirtyc && specc == 0u)) {
            printf("MISMATCH (GROUP BY phase 1, CR mode) delimiter %02x bytes", delim);
            for (int i = 0; i < 16; i++) printf(" %02x", b[i]);
            printf(": N %04x (expected %04x) D %04x (expected %04x) dirty %d spec %08x\n", n16c, enc, d16c, ed, (int)dirtyc, specc);
            exit(1);
        }
    }
    if (t16 != et || d16 != ed) {
        printf("MISMATCH delimiter %02x bytes", delim);
        for (int i = 0; i < 16; i++) printf(" %02x", b[i]);
        printf(": T %04x (expected %04x) D %04x (expected %04x)\n", t16, et, d16, ed);
        exit(1);
    }
    return 1;
}

int main() {
    const unsigned char delims[] = {',', ';', '|', ':', '~', '#'};
    long long n = 0;
    unsigned char b[16];
    for (unsigned char delim : delims) {
        // every pair of byte values in adjacent positions, at every position
        for (int pos = 0; pos < 15; pos++)
            for (int x = 0; x < 256; x++)
                for (int y = 0; y < 256; y += (pos % 4 == 3 ? 1 : 5)) {  // across a word boundary: every y
                    memset(b, 'a', 16);
                    b[pos] = (unsigned char)x;
                    b[pos + 1] = (unsigned char)y;
                    n += check(b, delim);
                }
        for (int r = 0; r < 300000; r++) {
            for (int i = 0; i < 16; i++) {
                const uint32_t k = rnd() % 8;
                b[i] = k == 0 ? delim : k == 1 ? '\n' : k == 2 ? (unsigned char)(rnd() % 0x30) : (unsigned char)rnd();
            }
            n += check(b, delim);
        }
    }
    printf("ok %lld chunks (%lld clean ones flagged dirty by the GROUP BY kernel's conservative test)\n", n, g_false_dirty);
    return 0;
}
