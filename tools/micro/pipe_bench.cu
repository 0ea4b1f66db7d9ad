// pipe_bench.cu — issue rate of the integer instructions the scan kernels are made of (sm_100a).
// Each test: 8 independent chains per thread, 32 warps per SM; prints warp-instructions per clock and SMSP.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_bench pipe_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITERS 2048

template <int OP>
__device__ __forceinline__ uint32_t step(uint32_t x, uint32_t k, uint32_t one) {
    uint32_t d;
    if (OP == 0) asm volatile("lop3.b32 %0, %1, %2, 0x5a5a5a5a, 0x96;" : "=r"(d) : "r"(x), "r"(k));
    else if (OP == 1) asm volatile("add.u32 %0, %1, 0x7f7f7f7f;" : "=r"(d) : "r"(x));
    else if (OP == 2) asm volatile("mad.lo.u32 %0, %1, %2, 0x7f7f7f7f;" : "=r"(d) : "r"(x), "r"(one));
    else if (OP == 3) asm volatile("mul.hi.u32 %0, %1, 0x02040810;" : "=r"(d) : "r"(x));
    else if (OP == 4) asm volatile("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(k), "r"(one));
    else if (OP == 5) asm volatile("prmt.b32 %0, %1, %2, 0x7740;" : "=r"(d) : "r"(x), "r"(k));
    else if (OP == 6) asm volatile("shf.r.wrap.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(k), "r"(one));
    else if (OP == 7) asm volatile("brev.b32 %0, %1;" : "=r"(d) : "r"(x));
    else if (OP == 8) asm volatile("bfind.u32 %0, %1;" : "=r"(d) : "r"(x));
    else if (OP == 9) asm volatile("popc.b32 %0, %1;" : "=r"(d) : "r"(x));
    else if (OP == 10) asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(x), "r"(k));
    else if (OP == 11) asm volatile("vabsdiff4.u32.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(k), "r"(one));
    else if (OP == 12) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(one), "r"(k));
    else if (OP == 13) {
        unsigned long long w;
        asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(w) : "r"(x), "r"(k), "l"(((unsigned long long)k << 32) | x));
        d = (uint32_t)w ^ (uint32_t)(w >> 32);
    } else if (OP == 14) asm volatile("redux.sync.add.u32 %0, %1, 0xffffffff;" : "=r"(d) : "r"(x));
    else if (OP == 15) asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(d) : "r"(x));
    else if (OP == 16) {
        asm volatile("{.reg .pred p; setp.ne.u32 p, %1, %2; vote.sync.ballot.b32 %0, p, 0xffffffff;}" : "=r"(d) : "r"(x), "r"(k));
    } else if (OP == 17) asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(d) : "r"(x), "r"(one));
    else if (OP == 18) asm volatile("match.any.sync.b32 %0, %1, 0xffffffff;" : "=r"(d) : "r"(x & 15u));
    else if (OP == 19) {
        // shared-memory atomic add, addresses spread over 16 words per warp
        asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(d) : "r"((x & 15u) * 4u), "r"(one) : "memory");
        d ^= x;
    } else d = x;
    return d;
}

// A, B: the two instruction kinds interleaved 1:1 (B = -1: A only)
template <int A, int B>
__global__ void __launch_bounds__(1024, 1) bench(uint32_t* out, long long* cycles, uint32_t one, uint32_t k) {
    __shared__ uint32_t sm_atom[64];
    if (threadIdx.x < 64) sm_atom[threadIdx.x] = 0;
    uint32_t x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) x[c] = threadIdx.x * 2654435761u + c;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
            x[c] = step<A>(x[c], k, one);
            if (B >= 0) x[c] = step<(B >= 0 ? B : 0)>(x[c], k, one);
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s ^= x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int A, int B>
static void run(const char* name, uint32_t* out, long long* cyc, int sms) {
    bench<A, B><<<sms, 1024>>>(out, cyc, 1u, 0x12345678u);
    bench<A, B><<<sms, 1024>>>(out, cyc, 1u, 0x12345678u);
    cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; i++) avg += (double)h[i];
    avg /= sms;
    const double insts = (double)ITERS * CHAINS * (B >= 0 ? 2 : 1) * 32 / 4;  // warp instructions per SMSP (8 warps each)
    printf("%-28s %.3f warp-inst/clk/SMSP\n", name, insts / avg);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t* out;
    long long* cyc;
    cudaMalloc(&out, (size_t)sms * 1024 * 4);
    cudaMalloc(&cyc, 256 * 8);
    run<0, -1>("LOP3", out, cyc, sms);
    run<1, -1>("add imm (VIADD?)", out, cyc, sms);
    run<10, -1>("add reg (IADD3?)", out, cyc, sms);
    run<2, -1>("IMAD r*r+imm", out, cyc, sms);
    run<12, -1>("IMAD r*r+r", out, cyc, sms);
    run<3, -1>("IMAD.HI", out, cyc, sms);
    run<4, -1>("IDP.4A", out, cyc, sms);
    run<5, -1>("PRMT", out, cyc, sms);
    run<6, -1>("SHF", out, cyc, sms);
    run<7, -1>("BREV", out, cyc, sms);
    run<8, -1>("FLO", out, cyc, sms);
    run<9, -1>("POPC", out, cyc, sms);
    run<11, -1>("VABSDIFF4", out, cyc, sms);
    run<13, -1>("IMAD.WIDE (+LOP3)", out, cyc, sms);
    run<14, -1>("REDUX.SUM", out, cyc, sms);
    run<15, -1>("REDUX.MIN", out, cyc, sms);
    run<16, -1>("ISETP + VOTE.ballot", out, cyc, sms);
    run<17, -1>("SHFL.IDX", out, cyc, sms);
    run<18, -1>("MATCH.ANY", out, cyc, sms);
    run<19, -1>("ATOMS.ADD (16 addresses)", out, cyc, sms);
    run<0, 14>("LOP3 + REDUX.SUM", out, cyc, sms);
    run<0, 1>("LOP3 + add imm", out, cyc, sms);
    run<0, 10>("LOP3 + add reg", out, cyc, sms);
    run<0, 2>("LOP3 + IMAD", out, cyc, sms);
    run<0, 3>("LOP3 + IMAD.HI", out, cyc, sms);
    run<0, 4>("LOP3 + IDP.4A", out, cyc, sms);
    run<0, 5>("LOP3 + PRMT", out, cyc, sms);
    run<0, 7>("LOP3 + BREV", out, cyc, sms);
    run<0, 8>("LOP3 + FLO", out, cyc, sms);
    run<0, 9>("LOP3 + POPC", out, cyc, sms);
    run<2, 3>("IMAD + IMAD.HI", out, cyc, sms);
    run<2, 4>("IMAD + IDP.4A", out, cyc, sms);
    run<1, 2>("add imm + IMAD", out, cyc, sms);
    return 0;
}
