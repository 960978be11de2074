// modmul_bench.cu — measures sustained 254-bit Montgomery products per second on the device
// (the ALU ceiling that bounds NTT / MSM / quotient).  Build & run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I webauthn-halo2_b200/csrc -o /tmp/mm tools/modmul_bench.cu && /tmp/mm
#include <cstdio>
#include <cuda_runtime.h>
#include "field.cuh"
using namespace zkw;

template <int ILP, class F>
__global__ void k_chain(uint32_t* out, int iters, uint32_t seed) {
    F x[ILP], y;
    for (int j = 0; j < ILP; j++) {
        x[j] = F::one();
        x[j].l[0] += threadIdx.x + j + seed;
    }
    y = F::one();
    y.l[1] ^= blockIdx.x + 12345u;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) x[j] = x[j] * y;
    }
    uint32_t acc = 0;
    for (int j = 0; j < ILP; j++) for (int l = 0; l < 8; l++) acc ^= x[j].l[l];
    if (acc == 0x12345678u) out[0] = acc;  // keep the chain alive
}

template <int ILP, class F>
static void run(const char* name, int blocks, int threads, int iters) {
    uint32_t* d;
    cudaMalloc(&d, 4);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    k_chain<ILP, F><<<blocks, threads>>>(d, 16, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k_chain<ILP, F><<<blocks, threads>>>(d, iters, 2);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    double muls = (double)blocks * threads * iters * ILP;
    printf("%-10s ILP=%d blocks=%d threads=%d: %.3f ms, %.2f Gmodmul/s\n", name, ILP, blocks, threads, ms, muls / ms / 1e6);
    cudaFree(d);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, %d MHz\n", p.name, p.multiProcessorCount, p.clockRate / 1000);
    int sm = p.multiProcessorCount;
    run<1, Fr>("Fr", sm * 8, 256, 2048);
    run<2, Fr>("Fr", sm * 8, 256, 2048);
    run<4, Fr>("Fr", sm * 4, 256, 2048);
    run<4, Fr>("Fr", sm * 4, 128, 2048);
    run<1, Fq>("Fq", sm * 8, 256, 2048);
    run<4, Fq>("Fq", sm * 4, 256, 2048);
    run<1, Fr>("Fr-1warp", 1, 32, 2048);  // single-warp latency per product
    return 0;
}
