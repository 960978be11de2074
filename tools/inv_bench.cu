// inv_bench.cu — device throughput of the two Fq inversions (Fermat a^(p-2) vs the binary GCD of
// inverse.cuh), in inversions/s and in Montgomery-product equivalents; also checks they agree on the device.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I webauthn-halo2_b200/csrc -o tools/bin/inv_bench tools/inv_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "inverse.cuh"
using namespace zkw;

template <int MODE>
__global__ void k_inv(uint32_t* out, int iters, uint32_t seed, uint32_t* mismatches) {
    Fq x = Fq::one();
    x.l[0] += threadIdx.x + seed;
    x.l[1] ^= blockIdx.x * 2654435761u;
    x.l[7] &= 0x0fffffffu;
    Fq acc = Fq::zero();
    for (int i = 0; i < iters; i++) {
        Fq r;
        if (MODE == 0) r = x.inv();
        else if (MODE == 1) r = fp_inv_bingcd(x);
        else {
            r = fp_inv_bingcd(x);
            Fq f = x.inv();
            if (!(r == f)) atomicAdd(mismatches, 1u);
        }
        acc = acc + r;
        x = x + r;
    }
    uint32_t a = 0;
    for (int l = 0; l < 8; l++) a ^= acc.l[l];
    if (a == 0x12345678u) out[0] = a;
}

template <int MODE>
static double run(const char* name, int blocks, int threads, int iters, uint32_t* d, uint32_t* mm) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k_inv<MODE><<<blocks, threads>>>(d, 1, 1, mm);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k_inv<MODE><<<blocks, threads>>>(d, iters, 2, mm);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double n = (double)blocks * threads * iters;
    printf("%-22s blocks=%d threads=%d iters=%d: %.3f ms, %.3f G inversions/s\n", name, blocks, threads, iters, ms, n / ms / 1e6);
    return n / ms / 1e6;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sm = p.multiProcessorCount;
    uint32_t *d, *mm;
    cudaMalloc(&d, 4); cudaMalloc(&mm, 4); cudaMemset(mm, 0, 4);
    const double f = run<0>("Fermat", sm * 8, 128, 8, d, mm);
    const double g = run<1>("binary GCD", sm * 8, 128, 32, d, mm);
    run<1>("binary GCD (4 warps/SM)", sm, 128, 32, d, mm);
    run<2>("both + compare", sm * 4, 128, 4, d, mm);
    uint32_t h = 0;
    cudaMemcpy(&h, mm, 4, cudaMemcpyDeviceToHost);
    printf("mismatches: %u\n", h);
    printf("at 68.5 G products/s: Fermat = %.0f product equivalents, binary GCD = %.0f\n", 68.5 / f, 68.5 / g);
    return h != 0;
}
