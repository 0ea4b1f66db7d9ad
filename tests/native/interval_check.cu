// Host-side exhaustive check of the comparison folding used by the scalar lean kernels (cqg_lean2.cuh):
// `value(field) <op> literal`, both exact decimals, must equal the modular interval test on the field's
// mantissa (lean2_interval) and, for fields of <= 4 digits, on its digit code (lean2_code).
// Built and run by tests/test_interval_logic.py (nvcc compiles it for the host; no GPU is used).
#include <cstdio>
#include <cstdlib>

#include "cqg_lean2.cuh"

using namespace cqg;

static long long p10(int k) {
    long long r = 1;
    while (k-- > 0) r *= 10;
    return r;
}

int main() {
    // literals as (mantissa, fraction digits); operators as in cq_gpu.h: compared through scaled integers
    const long long lits[][2] = {{0, 0}, {1, 0}, {25, 0}, {40, 0}, {80, 0}, {9999, 0}, {10000, 0}, {123456, 0}, {15, 1}, {5, 1},
                                 {125, 2}, {184, 2}, {1, 3}, {999, 3}, {2500, 2}, {33500, 3}, {7, 0}, {70, 1}, {1000000000ll, 0},
                                 {-1, 0}, {-25, 0}, {-15, 1}, {-125, 2}, {-9999, 0}, {-1, 3}, {-123456, 0}};
    const char* opname[] = {">", ">=", "<", "<=", "==", "!="};
    long long checked = 0;
    for (auto& lit : lits) {
        const long long cm = lit[0];
        const int cfd = (int)lit[1];
        for (int op = 0; op < 6; op++) {
            LeanLeaf L{};
            L.kind = 0;
            for (int fd = 0; fd < 4; fd++) {  // as make_lean_leaf (cqg_api.cu) fills A / LB
                const int K = fd > cfd ? fd : cfd;
                L.A[fd] = (uint32_t)p10(K - fd);
                L.LB[fd] = cm * p10(K - cfd) + (op == 1 ? -1 : op == 3 ? 1 : 0);
            }
            L.lop = (op == 0 || op == 1) ? 0 : (op == 2 || op == 3) ? 1 : op == 4 ? 2 : 3;
            for (int fd = 0; fd < 4; fd++) {
                uint32_t lo, width, clo, cwidth;
                lean2_interval(L, fd, lo, width, clo, cwidth);
                const int K = fd > cfd ? fd : cfd;
                const long long rhs = cm * p10(K - cfd);
                auto check = [&](uint32_t mant) {
                    const long long lhs = (long long)mant * p10(K - fd);
                    const bool want = op == 0 ? lhs > rhs : op == 1 ? lhs >= rhs : op == 2 ? lhs < rhs : op == 3 ? lhs <= rhs
                                    : op == 4 ? lhs == rhs : lhs != rhs;
                    const bool got = (uint32_t)(mant - lo) <= width;
                    if (got != want) {
                        printf("MISMATCH value: mant %u fd %d %s %lld/10^%d: want %d got %d (lo %u width %u)\n", mant, fd, opname[op], cm,
                               cfd, (int)want, (int)got, lo, width);
                        exit(1);
                    }
                    if (mant <= 9999u) {
                        const bool gotc = (uint32_t)(lean2_code(mant) - clo) <= cwidth;
                        if (gotc != want) {
                            printf("MISMATCH code: mant %u fd %d %s %lld/10^%d: want %d got %d (clo %08x cwidth %08x)\n", mant, fd, opname[op],
                                   cm, cfd, (int)want, (int)gotc, clo, cwidth);
                            exit(1);
                        }
                    }
                    checked++;
                };
                for (uint32_t m = 0; m <= 12000u; m++) check(m);
                for (uint32_t m = 12000u; m < 10000000u; m += 9973u) check(m);
                check(9999999u);
                // a field with a leading '-': value = -(mant / 10^fd), the interval 8 bytes behind (lean2_interval_neg)
                uint32_t nlo, nwidth;
                lean2_interval_neg(L, fd, nlo, nwidth);
                auto check_neg = [&](uint32_t mant) {
                    const long long lhs = -(long long)mant * p10(K - fd);
                    const bool want = op == 0 ? lhs > rhs : op == 1 ? lhs >= rhs : op == 2 ? lhs < rhs : op == 3 ? lhs <= rhs
                                    : op == 4 ? lhs == rhs : lhs != rhs;
                    const bool got = (uint32_t)(mant - nlo) <= nwidth;
                    if (got != want) {
                        printf("MISMATCH negative value: -mant %u fd %d %s %lld/10^%d: want %d got %d (lo %u width %u)\n", mant, fd, opname[op],
                               cm, cfd, (int)want, (int)got, nlo, nwidth);
                        exit(1);
                    }
                    checked++;
                };
                for (uint32_t m = 0; m <= 12000u; m++) check_neg(m);
                for (uint32_t m = 12000u; m < 1000000u; m += 997u) check_neg(m);
                check_neg(999999u);
            }
        }
    }
    printf("ok %lld comparisons\n", checked);
    return 0;
}
