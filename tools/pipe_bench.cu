// pipe_bench.cu — issue rates of the integer / fp64 instructions a 254-bit Montgomery product can be
// built from, in lane-operations per clock per SM (development aid; explains the ALU ceiling in DESIGN.md).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipe tools/pipe_bench.cu && /tmp/pipe
// Every multiply takes a LOOP-VARIANT operand (the low word of the neighbouring chain's accumulator), otherwise
// ptxas hoists the product out of the loop and the "multiply" test times additions (the round-1 version did exactly
// that for the plain mad.wide line).  tools/pipe_bench_sass.sh dumps the loop bodies' SASS so that the mnemonic each
// line really measures is on record (profiles/r2_pipe_sass.txt).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define U 8

enum Op { MAD_WIDE, MAD_LOHI_CC, MAD_LO, MAD_HI, DFMA, IADD3_CC, LOP3, MAD_WIDE_PLUS_ALU, DFMA_PLUS_MADWIDE, MAD_CHAIN4 };

template <int OP>
__global__ void k(uint32_t* out, uint32_t seed) {
    uint32_t a = threadIdx.x * 2654435761u + seed, b = blockIdx.x * 40503u + 77u;
    uint64_t w[U];
    uint32_t lo[U], hi[U];
    double d[U];
    double da = 1.0 + (double)(threadIdx.x & 7) * 1e-9, db = 1.0 - 1e-9;
#pragma unroll
    for (int j = 0; j < U; j++) { w[j] = j + seed + ((uint64_t)a << 7); lo[j] = j + a; hi[j] = j ^ b ^ a; d[j] = (double)j; }   // thread-variant: keeps the work off the uniform datapath
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < U; j++) {
            const uint32_t va = (uint32_t)w[(j + 1) % U], vl = lo[(j + 1) % U];   // loop-variant multiplicands
            if (OP == MAD_WIDE) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[j]) : "r"(va), "r"(b));
            if (OP == MAD_LOHI_CC)
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[j]), "+r"(hi[j]) : "r"(vl), "r"(b));
            if (OP == MAD_LO) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[j]) : "r"(vl), "r"(b));
            if (OP == MAD_HI) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(lo[j]) : "r"(vl), "r"(b));
            if (OP == DFMA) asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d[j]) : "d"(da), "d"(db));
            if (OP == IADD3_CC) asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(lo[j]), "+r"(hi[j]) : "r"(a), "r"(b));
            if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(lo[j]) : "r"(a), "r"(hi[j]));
            if (OP == MAD_WIDE_PLUS_ALU) {
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[j]) : "r"(va), "r"(b));
                asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(lo[j]), "+r"(hi[j]) : "r"(a), "r"(b));
            }
            if (OP == DFMA_PLUS_MADWIDE) {
                asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d[j]) : "d"(da), "d"(db));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[j]) : "r"(va), "r"(b));
            }
        }
    }
    if (OP == MAD_CHAIN4) {
        for (int i = 0; i < ITERS; i++) {
#pragma unroll
            for (int j = 0; j < U; j += 4)
                asm volatile("mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\t"
                             "madc.lo.cc.u32 %2, %8, %9, %2;\n\tmadc.hi.cc.u32 %3, %8, %9, %3;\n\t"
                             "madc.lo.cc.u32 %4, %8, %9, %4;\n\tmadc.hi.cc.u32 %5, %8, %9, %5;\n\t"
                             "madc.lo.cc.u32 %6, %8, %9, %6;\n\tmadc.hi.u32 %7, %8, %9, %7;"
                             : "+r"(lo[j]), "+r"(hi[j]), "+r"(lo[j + 1]), "+r"(hi[j + 1]), "+r"(lo[j + 2]), "+r"(hi[j + 2]), "+r"(lo[j + 3]), "+r"(hi[j + 3])
                             : "r"(lo[(j + 4) % U]), "r"(b));
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < U; j++) acc ^= (uint32_t)w[j] ^ (uint32_t)(w[j] >> 32) ^ lo[j] ^ hi[j] ^ (uint32_t)__double_as_longlong(d[j]);
    if (acc == 0x12345678u) out[0] = acc;
}

template <int OP>
static void run(const char* name, int per_iter, int sm, double mhz) {
    uint32_t* d;
    cudaMalloc(&d, 4);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int blocks = sm * 4, threads = 256;
    k<OP><<<blocks, threads>>>(d, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<OP><<<blocks, threads>>>(d, 2);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double ops = (double)blocks * threads * ITERS * U * per_iter;
    printf("%-28s %8.3f ms  %7.1f lane-instr/clk/SM (at %.0f MHz)\n", name, ms, ops / (ms * 1e-3) / (mhz * 1e6) / sm, mhz);
    cudaFree(d);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sm = p.multiProcessorCount;
    const double mhz = p.clockRate / 1000.0;
    printf("%s, %d SMs, %.0f MHz (nominal; rates assume this clock)\n", p.name, sm, mhz);
    run<MAD_WIDE>("mad.wide.u32 (IMAD.WIDE)", 1, sm, mhz);
    run<MAD_LOHI_CC>("mad.lo.cc+madc.hi.cc pair", 1, sm, mhz);
    run<MAD_LO>("mad.lo.u32 (IMAD)", 1, sm, mhz);
    run<MAD_HI>("mad.hi.u32 (IMAD.HI)", 1, sm, mhz);
    run<DFMA>("fma.rz.f64 (DFMA)", 1, sm, mhz);
    run<IADD3_CC>("add.cc+addc pair (2 instr)", 2, sm, mhz);
    run<LOP3>("lop3", 1, sm, mhz);
    run<MAD_WIDE_PLUS_ALU>("mad.wide + add.cc/addc (3)", 3, sm, mhz);
    run<DFMA_PLUS_MADWIDE>("dfma + mad.wide (2)", 2, sm, mhz);
    run<MAD_CHAIN4>("4-pair carry chain (.X), per pair", 1, sm, mhz);
    return 0;
}
