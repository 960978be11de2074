// field.cuh — 254-bit Montgomery prime fields (BN254 Fr and Fq) in registers, for sm_100a.
//
// Device-side replacement for halo2curves::bn256::{Fr, Fq} (imported by the reference at
// halo2-circuits/src/ecc/ecdsa_p256.rs:27; moduli = f_q / f_p of proving-server/P256Verifier.yul:17-18).
// Memory format is halo2curves' own: [u64;4] little-endian limbs, Montgomery form with R = 2^256,
// fully reduced — read here as 8 x u32 (same bytes) with two 128-bit loads per element.
//
// The product is a word-serial Montgomery multiplication over 32-bit limbs built from
// mad.lo.cc / madc.hi.cc carry chains.  Partial products of even- and odd-indexed multiplicand
// limbs are kept in two accumulators (A at limb offset 0, B at limb offset 1) so each row is one
// uninterrupted carry chain; the per-row division by 2^32 swaps the roles of the accumulators
// (A' = B + A[1], B' = A >> 64) and is folded into the next row's chain.  No tensor cores: this is
// integer modular arithmetic.
#pragma once
#include <stdint.h>
#include <string.h>

namespace zkw {

struct FrParams {
    // r = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
    static __host__ __device__ constexpr uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return m[i];
    }
    static __host__ __device__ constexpr uint32_t one(int i) {  // R mod r
        constexpr uint32_t m[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return m[i];
    }
    static __host__ __device__ constexpr uint32_t r2(int i) {  // R^2 mod r
        constexpr uint32_t m[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
        return m[i];
    }
    static constexpr uint32_t INV = 0xefffffffu;  // -r^-1 mod 2^32
};

struct FqParams {
    // p = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
    static __host__ __device__ constexpr uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return m[i];
    }
    static __host__ __device__ constexpr uint32_t one(int i) {  // R mod p
        constexpr uint32_t m[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return m[i];
    }
    static __host__ __device__ constexpr uint32_t r2(int i) {  // R^2 mod p
        constexpr uint32_t m[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
        return m[i];
    }
    static constexpr uint32_t INV = 0xe4866389u;  // -p^-1 mod 2^32
};

// c[0..7] = {lo,hi} of x0*b, x1*b, x2*b, x3*b (no carries: the four products do not overlap)
__device__ __forceinline__ void mul_pairs(uint32_t c[9], uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t b) {
    asm("mul.lo.u32 %0, %8, %12;\n\t"
        "mul.hi.u32 %1, %8, %12;\n\t"
        "mul.lo.u32 %2, %9, %12;\n\t"
        "mul.hi.u32 %3, %9, %12;\n\t"
        "mul.lo.u32 %4, %10, %12;\n\t"
        "mul.hi.u32 %5, %10, %12;\n\t"
        "mul.lo.u32 %6, %11, %12;\n\t"
        "mul.hi.u32 %7, %11, %12;"
        : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]), "=r"(c[4]), "=r"(c[5]), "=r"(c[6]), "=r"(c[7])
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b));
    c[8] = 0;
}

// c[0..8] += x0*b + x1*b*2^64 + x2*b*2^128 + x3*b*2^192 as one carry chain
__device__ __forceinline__ void mad_pairs(uint32_t c[9], uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t b) {
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(c[6]), "+r"(c[7]), "+r"(c[8])
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b));
}

// Row step of the two-accumulator scheme after a division by 2^32:
//   nA0 = B0 + A1 (carry out), nB[k] = A[k+2] + (x_k * b pairs) + carry, k = 0..7 (A[9] == 0)
__device__ __forceinline__ void shift_mad_pairs(uint32_t& nA0, uint32_t nB[9], uint32_t B0, const uint32_t A[9],
                                                uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t b) {
    asm("add.cc.u32 %0, %9, %10;\n\t"
        "madc.lo.cc.u32 %1, %18, %22, %11;\n\t"
        "madc.hi.cc.u32 %2, %18, %22, %12;\n\t"
        "madc.lo.cc.u32 %3, %19, %22, %13;\n\t"
        "madc.hi.cc.u32 %4, %19, %22, %14;\n\t"
        "madc.lo.cc.u32 %5, %20, %22, %15;\n\t"
        "madc.hi.cc.u32 %6, %20, %22, %16;\n\t"
        "madc.lo.cc.u32 %7, %21, %22, %17;\n\t"
        "madc.hi.u32 %8, %21, %22, 0;"
        : "=r"(nA0), "=r"(nB[0]), "=r"(nB[1]), "=r"(nB[2]), "=r"(nB[3]), "=r"(nB[4]), "=r"(nB[5]), "=r"(nB[6]), "=r"(nB[7])
        : "r"(B0), "r"(A[1]), "r"(A[2]), "r"(A[3]), "r"(A[4]), "r"(A[5]), "r"(A[6]), "r"(A[7]), "r"(A[8]),
          "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b));
    nB[8] = 0;
}


// ---- irregular rows of the squaring (Fp::sqr_lazy): the same two chains with their first S pairs absent ------------------
// c[2S..8] += x_S*b*2^(64 S) + ... + x_3*b*2^192
__device__ __forceinline__ void mad_pairs_s1(uint32_t c[9], uint32_t x1, uint32_t x2, uint32_t x3, uint32_t b) {
    asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\t"
        "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
        "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
        "madc.lo.cc.u32 %4, %9, %10, %4;\n\t"
        "madc.hi.cc.u32 %5, %9, %10, %5;\n\t"
        "addc.u32 %6, %6, 0;"
        : "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(c[6]), "+r"(c[7]), "+r"(c[8])
        : "r"(x1), "r"(x2), "r"(x3), "r"(b));
}
__device__ __forceinline__ void mad_pairs_s2(uint32_t c[9], uint32_t x2, uint32_t x3, uint32_t b) {
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+r"(c[4]), "+r"(c[5]), "+r"(c[6]), "+r"(c[7]), "+r"(c[8])
        : "r"(x2), "r"(x3), "r"(b));
}
__device__ __forceinline__ void mad_pairs_s3(uint32_t c[9], uint32_t x3, uint32_t b) {
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32 %2, %2, 0;"
        : "+r"(c[6]), "+r"(c[7]), "+r"(c[8])
        : "r"(x3), "r"(b));
}
// shift_mad_pairs with the first S product pairs absent: those limbs only take the shifted accumulator and the carry
__device__ __forceinline__ void shift_mad_pairs_s1(uint32_t& nA0, uint32_t nB[9], uint32_t B0, const uint32_t A[9],
                                                   uint32_t x1, uint32_t x2, uint32_t x3, uint32_t b) {
    asm("add.cc.u32 %0, %9, %10;\n\t"
        "addc.cc.u32 %1, %11, 0;\n\t"
        "addc.cc.u32 %2, %12, 0;\n\t"
        "madc.lo.cc.u32 %3, %18, %21, %13;\n\t"
        "madc.hi.cc.u32 %4, %18, %21, %14;\n\t"
        "madc.lo.cc.u32 %5, %19, %21, %15;\n\t"
        "madc.hi.cc.u32 %6, %19, %21, %16;\n\t"
        "madc.lo.cc.u32 %7, %20, %21, %17;\n\t"
        "madc.hi.u32 %8, %20, %21, 0;"
        : "=r"(nA0), "=r"(nB[0]), "=r"(nB[1]), "=r"(nB[2]), "=r"(nB[3]), "=r"(nB[4]), "=r"(nB[5]), "=r"(nB[6]), "=r"(nB[7])
        : "r"(B0), "r"(A[1]), "r"(A[2]), "r"(A[3]), "r"(A[4]), "r"(A[5]), "r"(A[6]), "r"(A[7]), "r"(A[8]),
          "r"(x1), "r"(x2), "r"(x3), "r"(b));
    nB[8] = 0;
}
__device__ __forceinline__ void shift_mad_pairs_s2(uint32_t& nA0, uint32_t nB[9], uint32_t B0, const uint32_t A[9],
                                                   uint32_t x2, uint32_t x3, uint32_t b) {
    asm("add.cc.u32 %0, %9, %10;\n\t"
        "addc.cc.u32 %1, %11, 0;\n\t"
        "addc.cc.u32 %2, %12, 0;\n\t"
        "addc.cc.u32 %3, %13, 0;\n\t"
        "addc.cc.u32 %4, %14, 0;\n\t"
        "madc.lo.cc.u32 %5, %18, %20, %15;\n\t"
        "madc.hi.cc.u32 %6, %18, %20, %16;\n\t"
        "madc.lo.cc.u32 %7, %19, %20, %17;\n\t"
        "madc.hi.u32 %8, %19, %20, 0;"
        : "=r"(nA0), "=r"(nB[0]), "=r"(nB[1]), "=r"(nB[2]), "=r"(nB[3]), "=r"(nB[4]), "=r"(nB[5]), "=r"(nB[6]), "=r"(nB[7])
        : "r"(B0), "r"(A[1]), "r"(A[2]), "r"(A[3]), "r"(A[4]), "r"(A[5]), "r"(A[6]), "r"(A[7]), "r"(A[8]),
          "r"(x2), "r"(x3), "r"(b));
    nB[8] = 0;
}
__device__ __forceinline__ void shift_mad_pairs_s3(uint32_t& nA0, uint32_t nB[9], uint32_t B0, const uint32_t A[9], uint32_t x3, uint32_t b) {
    asm("add.cc.u32 %0, %9, %10;\n\t"
        "addc.cc.u32 %1, %11, 0;\n\t"
        "addc.cc.u32 %2, %12, 0;\n\t"
        "addc.cc.u32 %3, %13, 0;\n\t"
        "addc.cc.u32 %4, %14, 0;\n\t"
        "addc.cc.u32 %5, %15, 0;\n\t"
        "addc.cc.u32 %6, %16, 0;\n\t"
        "madc.lo.cc.u32 %7, %18, %19, %17;\n\t"
        "madc.hi.u32 %8, %18, %19, 0;"
        : "=r"(nA0), "=r"(nB[0]), "=r"(nB[1]), "=r"(nB[2]), "=r"(nB[3]), "=r"(nB[4]), "=r"(nB[5]), "=r"(nB[6]), "=r"(nB[7])
        : "r"(B0), "r"(A[1]), "r"(A[2]), "r"(A[3]), "r"(A[4]), "r"(A[5]), "r"(A[6]), "r"(A[7]), "r"(A[8]),
          "r"(x3), "r"(b));
    nB[8] = 0;
}

template <class P>
struct Fp {
    uint32_t l[8];

    __host__ __device__ __forceinline__ static Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = 0;
        return r;
    }
    __host__ __device__ __forceinline__ static Fp one() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = P::one(i);
        return r;
    }
    __host__ __device__ __forceinline__ static Fp r2() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = P::r2(i);
        return r;
    }
    // two 128-bit loads / stores (element = 32 B, 16-B aligned at minimum)
    __host__ __device__ __forceinline__ static Fp load(const void* p) {
        const uint4* q = reinterpret_cast<const uint4*>(p);
        uint4 lo = q[0], hi = q[1];
        Fp r;
        r.l[0] = lo.x; r.l[1] = lo.y; r.l[2] = lo.z; r.l[3] = lo.w;
        r.l[4] = hi.x; r.l[5] = hi.y; r.l[6] = hi.z; r.l[7] = hi.w;
        return r;
    }
    __device__ __forceinline__ static Fp load_nc(const void* p) {  // read-only path
        const uint4* q = reinterpret_cast<const uint4*>(p);
        uint4 lo = __ldg(q), hi = __ldg(q + 1);
        Fp r;
        r.l[0] = lo.x; r.l[1] = lo.y; r.l[2] = lo.z; r.l[3] = lo.w;
        r.l[4] = hi.x; r.l[5] = hi.y; r.l[6] = hi.z; r.l[7] = hi.w;
        return r;
    }
    __device__ __forceinline__ static Fp load_cg(const void* p) {  // L2 only: data another CTA of the same launch just wrote
        const uint4* q = reinterpret_cast<const uint4*>(p);
        uint4 lo = __ldcg(q), hi = __ldcg(q + 1);
        Fp r;
        r.l[0] = lo.x; r.l[1] = lo.y; r.l[2] = lo.z; r.l[3] = lo.w;
        r.l[4] = hi.x; r.l[5] = hi.y; r.l[6] = hi.z; r.l[7] = hi.w;
        return r;
    }
    __host__ __device__ __forceinline__ void store(void* p) const {
        uint4* q = reinterpret_cast<uint4*>(p);
        q[0] = make_uint4(l[0], l[1], l[2], l[3]);
        q[1] = make_uint4(l[4], l[5], l[6], l[7]);
    }
    __host__ __device__ __forceinline__ bool is_zero() const {
        return (l[0] | l[1] | l[2] | l[3] | l[4] | l[5] | l[6] | l[7]) == 0;
    }
    __host__ __device__ __forceinline__ bool operator==(const Fp& o) const {
        uint32_t d = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) d |= l[i] ^ o.l[i];
        return d == 0;
    }
    __host__ __device__ __forceinline__ bool operator!=(const Fp& o) const { return !(*this == o); }

    // t = x - mod; returns borrow (1 if x < mod)
    __host__ __device__ __forceinline__ static uint32_t sub_mod(uint32_t t[8], const uint32_t x[8]) {
        uint32_t borrow;
#ifndef __CUDA_ARCH__
        int64_t c = 0;
        for (int i = 0; i < 8; i++) { c += (int64_t)x[i] - (int64_t)P::mod(i); t[i] = (uint32_t)c; c >>= 32; }
        borrow = (uint32_t)c;
#else
        asm("sub.cc.u32 %0, %9, %17;\n\t"
            "subc.cc.u32 %1, %10, %18;\n\t"
            "subc.cc.u32 %2, %11, %19;\n\t"
            "subc.cc.u32 %3, %12, %20;\n\t"
            "subc.cc.u32 %4, %13, %21;\n\t"
            "subc.cc.u32 %5, %14, %22;\n\t"
            "subc.cc.u32 %6, %15, %23;\n\t"
            "subc.cc.u32 %7, %16, %24;\n\t"
            "subc.u32 %8, 0, 0;"
            : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(borrow)
            : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]),
              "r"(P::mod(0)), "r"(P::mod(1)), "r"(P::mod(2)), "r"(P::mod(3)), "r"(P::mod(4)), "r"(P::mod(5)), "r"(P::mod(6)), "r"(P::mod(7)));
#endif
        return borrow;  // 0xffffffff if borrowed, else 0
    }
    // x < 2*mod  ->  x mod mod
    __host__ __device__ __forceinline__ void reduce_once() {
        uint32_t t[8];
        uint32_t borrow = sub_mod(t, l);
#pragma unroll
        for (int i = 0; i < 8; i++) l[i] = borrow ? l[i] : t[i];
    }

    __host__ __device__ __forceinline__ friend Fp operator+(const Fp& a, const Fp& b) {
        Fp r;
#ifndef __CUDA_ARCH__
        uint64_t c = 0;
        for (int i = 0; i < 8; i++) { c += (uint64_t)a.l[i] + b.l[i]; r.l[i] = (uint32_t)c; c >>= 32; }
#else
        asm("add.cc.u32 %0, %8, %16;\n\t"
            "addc.cc.u32 %1, %9, %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32 %7, %15, %23;"
            : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7])
            : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
              "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
#endif
        r.reduce_once();  // a,b < mod < 2^254: the sum cannot carry out of 256 bits
        return r;
    }
    __host__ __device__ __forceinline__ friend Fp operator-(const Fp& a, const Fp& b) {
        Fp r;
        uint32_t borrow;
#ifndef __CUDA_ARCH__
        int64_t c = 0;
        for (int i = 0; i < 8; i++) { c += (int64_t)a.l[i] - (int64_t)b.l[i]; r.l[i] = (uint32_t)c; c >>= 32; }
        borrow = (uint32_t)c;
        uint64_t d = 0;
        for (int i = 0; i < 8; i++) { d += (uint64_t)r.l[i] + (P::mod(i) & borrow); r.l[i] = (uint32_t)d; d >>= 32; }
#else
        asm("sub.cc.u32 %0, %9, %17;\n\t"
            "subc.cc.u32 %1, %10, %18;\n\t"
            "subc.cc.u32 %2, %11, %19;\n\t"
            "subc.cc.u32 %3, %12, %20;\n\t"
            "subc.cc.u32 %4, %13, %21;\n\t"
            "subc.cc.u32 %5, %14, %22;\n\t"
            "subc.cc.u32 %6, %15, %23;\n\t"
            "subc.cc.u32 %7, %16, %24;\n\t"
            "subc.u32 %8, 0, 0;"
            : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7]), "=r"(borrow)
            : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
              "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
        // add back mod & borrow-mask
        asm("add.cc.u32 %0, %0, %8;\n\t"
            "addc.cc.u32 %1, %1, %9;\n\t"
            "addc.cc.u32 %2, %2, %10;\n\t"
            "addc.cc.u32 %3, %3, %11;\n\t"
            "addc.cc.u32 %4, %4, %12;\n\t"
            "addc.cc.u32 %5, %5, %13;\n\t"
            "addc.cc.u32 %6, %6, %14;\n\t"
            "addc.u32 %7, %7, %15;"
            : "+r"(r.l[0]), "+r"(r.l[1]), "+r"(r.l[2]), "+r"(r.l[3]), "+r"(r.l[4]), "+r"(r.l[5]), "+r"(r.l[6]), "+r"(r.l[7])
            : "r"(P::mod(0) & borrow), "r"(P::mod(1) & borrow), "r"(P::mod(2) & borrow), "r"(P::mod(3) & borrow),
              "r"(P::mod(4) & borrow), "r"(P::mod(5) & borrow), "r"(P::mod(6) & borrow), "r"(P::mod(7) & borrow));
#endif
        return r;
    }
    __host__ __device__ __forceinline__ Fp neg() const { return zero() - *this; }
    __host__ __device__ __forceinline__ Fp dbl() const { return *this + *this; }

    // Montgomery product a*b*R^-1 mod m, fully reduced.
    __host__ __device__ __forceinline__ friend Fp operator*(const Fp& a, const Fp& b) {
#ifndef __CUDA_ARCH__
        // host path: word-serial (CIOS) Montgomery product over four 64-bit limbs with 128-bit partial products.
        // The host runs the last ~2c point operations and one inversion of every MSM and the transcript-side
        // field arithmetic of the prover, between the device's rounds - it is on the proof's critical path.
        typedef unsigned __int128 u128;
        uint64_t x[4], y[4], md[4];
        memcpy(x, a.l, 32);   // eight little-endian u32 limbs are four little-endian u64 limbs (x86-64 host)
        memcpy(y, b.l, 32);
        for (int i = 0; i < 4; i++) md[i] = (uint64_t)P::mod(2 * i) | ((uint64_t)P::mod(2 * i + 1) << 32);
        // -m^-1 mod 2^64 from the 32-bit constant: one Newton step on m^-1 mod 2^32
        const uint64_t minv32 = (uint32_t)(0u - P::INV);
        const uint64_t inv64 = 0 - minv32 * (2 - md[0] * minv32);
        uint64_t t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; i++) {
            u128 c = 0;
            for (int j = 0; j < 4; j++) { c += (u128)x[j] * y[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
            c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
            const uint64_t m = t[0] * inv64;
            c = (u128)m * md[0] + t[0]; c >>= 64;
            for (int j = 1; j < 4; j++) { c += (u128)m * md[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
            c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
        }
        Fp r;
        memcpy(r.l, t, 32);
        r.reduce_once();   // a, b < m < 2^254: the result is below 2m and t[4] is zero
        return r;
#else
        uint32_t A[9], B[9];
        mul_pairs(A, a.l[0], a.l[2], a.l[4], a.l[6], b.l[0]);
        mul_pairs(B, a.l[1], a.l[3], a.l[5], a.l[7], b.l[0]);
        uint32_t m = A[0] * P::INV;
        mad_pairs(A, P::mod(0), P::mod(2), P::mod(4), P::mod(6), m);
        mad_pairs(B, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
#pragma unroll
        for (int i = 1; i < 8; i++) {
            uint32_t nA[9], nB[9];
            shift_mad_pairs(nA[0], nB, B[0], A, a.l[1], a.l[3], a.l[5], a.l[7], b.l[i]);
#pragma unroll
            for (int k = 1; k < 9; k++) nA[k] = B[k];
            mad_pairs(nA, a.l[0], a.l[2], a.l[4], a.l[6], b.l[i]);
            m = nA[0] * P::INV;
            mad_pairs(nA, P::mod(0), P::mod(2), P::mod(4), P::mod(6), m);
            mad_pairs(nB, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
#pragma unroll
            for (int k = 0; k < 9; k++) { A[k] = nA[k]; B[k] = nB[k]; }
        }
        Fp r;
        asm("add.cc.u32 %0, %8, %16;\n\t"
            "addc.cc.u32 %1, %9, %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32 %7, %15, %23;"
            : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7])
            : "r"(A[1]), "r"(A[2]), "r"(A[3]), "r"(A[4]), "r"(A[5]), "r"(A[6]), "r"(A[7]), "r"(A[8]),
              "r"(B[0]), "r"(B[1]), "r"(B[2]), "r"(B[3]), "r"(B[4]), "r"(B[5]), "r"(B[6]), "r"(B[7]));
        r.reduce_once();
        return r;
#endif
    }
    // canonical square: on the device the dedicated squaring below (bit-identical to the product, 100 wide multiplies)
    __host__ __device__ __forceinline__ Fp sqr() const {
#ifdef __CUDA_ARCH__
        return sqr_lazy(*this).normalized();
#else
        return *this * *this;
#endif
    }

    // ---- lazily reduced arithmetic: values in [0, 2m) ------------------------------------------------------
    // 4m < 2^256 for both BN254 moduli, so a Montgomery product of two values below 2m is below 2m WITHOUT the
    // final conditional subtraction (a*b/R + m < 4m^2/R + m < 2m), and sums of two such values still fit 256
    // bits.  The bucket accumulation keeps its running point in this form: on this machine every instruction
    // costs dispatch cycles (IMAD.WIDE 4, IMAD 2, the rest 1 - the model that fits the measured 548 cycles per
    // product and 5770 per mixed addition), so the 17 instructions of a reduce_once are not hidden behind the
    // multiplies.  Values are brought back below m (normalized()) before they leave the kernel.
    static __host__ __device__ constexpr uint32_t mod2(int i) {   // limb i of 2m
        return (P::mod(i) << 1) | (i ? (P::mod(i - 1) >> 31) : 0u);
    }
    __host__ __device__ __forceinline__ Fp normalized() const {   // [0, 2m) -> [0, m)
        Fp r = *this;
        r.reduce_once();
        return r;
    }
    __host__ __device__ __forceinline__ bool is_zero_lazy() const {   // value == 0 mod m, for values in [0, 2m)
        uint32_t d = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) d |= l[i] ^ P::mod(i);
        return is_zero() || d == 0;
    }
    // a * b * R^-1 mod m for a, b < 2m; result < 2m
    __device__ __forceinline__ static Fp mul_lazy(const Fp& a, const Fp& b) {
#ifndef __CUDA_ARCH__
        return a.normalized() * b.normalized();
#else
        uint32_t A[9], B[9];
        mul_pairs(A, a.l[0], a.l[2], a.l[4], a.l[6], b.l[0]);
        mul_pairs(B, a.l[1], a.l[3], a.l[5], a.l[7], b.l[0]);
        uint32_t m = A[0] * P::INV;
        mad_pairs(A, P::mod(0), P::mod(2), P::mod(4), P::mod(6), m);
        mad_pairs(B, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
#pragma unroll
        for (int i = 1; i < 8; i++) {
            uint32_t nA[9], nB[9];
            shift_mad_pairs(nA[0], nB, B[0], A, a.l[1], a.l[3], a.l[5], a.l[7], b.l[i]);
#pragma unroll
            for (int k = 1; k < 9; k++) nA[k] = B[k];
            mad_pairs(nA, a.l[0], a.l[2], a.l[4], a.l[6], b.l[i]);
            m = nA[0] * P::INV;
            mad_pairs(nA, P::mod(0), P::mod(2), P::mod(4), P::mod(6), m);
            mad_pairs(nB, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
#pragma unroll
            for (int k = 0; k < 9; k++) { A[k] = nA[k]; B[k] = nB[k]; }
        }
        Fp r;
        asm("add.cc.u32 %0, %8, %16;\n\t"
            "addc.cc.u32 %1, %9, %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32 %7, %15, %23;"
            : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7])
            : "r"(A[1]), "r"(A[2]), "r"(A[3]), "r"(A[4]), "r"(A[5]), "r"(A[6]), "r"(A[7]), "r"(A[8]),
              "r"(B[0]), "r"(B[1]), "r"(B[2]), "r"(B[3]), "r"(B[4]), "r"(B[5]), "r"(B[6]), "r"(B[7]));
        return r;
#endif
    }
    // a * a * R^-1 mod m for a < 2m; result < 2m - bit-identical to mul_lazy(a, a), with 100 wide multiplies instead of 128:
    // row i of the interleaved product only adds a_i*a_i and a_i * 2*(a div 2^(32(i+1))) (the pairs j < i were added, doubled, in
    // row j), i.e. the two carry chains lose their first ceil(i/2) / floor(i/2) pairs.  The doubled upper part's limbs are those
    // of d = 2a, except its lowest one (j = i + 1), which must not carry the bit shifted out of a_i: e_j = (a_j << 1) mod 2^32.  The running limb that fixes m in row i
    // only depends on pairs of total weight <= i, all of which have been added by then, so every m - and therefore the
    // result - equals the full product's.  2a < 4m < 2^256 fits eight limbs, and a_i * 2a + m * p < 2^288 keeps the nine-limb
    // accumulators from overflowing.
    __device__ __forceinline__ static Fp sqr_lazy(const Fp& a) {
#ifndef __CUDA_ARCH__
        return a.normalized() * a.normalized();
#else
        uint32_t d[8];   // 2a
        asm("add.cc.u32 %0, %8, %8;\n\t"
            "addc.cc.u32 %1, %9, %9;\n\t"
            "addc.cc.u32 %2, %10, %10;\n\t"
            "addc.cc.u32 %3, %11, %11;\n\t"
            "addc.cc.u32 %4, %12, %12;\n\t"
            "addc.cc.u32 %5, %13, %13;\n\t"
            "addc.cc.u32 %6, %14, %14;\n\t"
            "addc.u32 %7, %15, %15;"
            : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7])
            : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]));
        uint32_t e[8];   // e_j = 2 a_j mod 2^32 (d_j without the carry from limb j - 1)
#pragma unroll
        for (int j = 1; j < 8; j++) e[j] = a.l[j] << 1;
        uint32_t A[9], B[9], nA[9], nB[9], m;
#define ZKW_SQR_REDUCE(AA, BB)                                               \
        m = AA[0] * P::INV;                                                  \
        mad_pairs(AA, P::mod(0), P::mod(2), P::mod(4), P::mod(6), m);        \
        mad_pairs(BB, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
#define ZKW_SQR_NEXT()                                                       \
        _Pragma("unroll") for (int k = 0; k < 9; k++) { A[k] = nA[k]; B[k] = nB[k]; }
#define ZKW_SQR_COPY_B()                                                     \
        _Pragma("unroll") for (int k = 1; k < 9; k++) nA[k] = B[k];
        // row 0: a0*a0, a0*e1, a0*d2 .. a0*d7
        mul_pairs(A, a.l[0], d[2], d[4], d[6], a.l[0]);
        mul_pairs(B, e[1], d[3], d[5], d[7], a.l[0]);
        ZKW_SQR_REDUCE(A, B)
        // row 1: odd chain a1*a1, d3, d5, d7; even chain d2, d4, d6 (pair 0 absent)
        shift_mad_pairs(nA[0], nB, B[0], A, a.l[1], d[3], d[5], d[7], a.l[1]);
        ZKW_SQR_COPY_B()
        mad_pairs_s1(nA, e[2], d[4], d[6], a.l[1]);
        ZKW_SQR_REDUCE(nA, nB)
        ZKW_SQR_NEXT()
        // row 2: odd chain d3, d5, d7 (pair 0 absent); even chain a2*a2, d4, d6
        shift_mad_pairs_s1(nA[0], nB, B[0], A, e[3], d[5], d[7], a.l[2]);
        ZKW_SQR_COPY_B()
        mad_pairs_s1(nA, a.l[2], d[4], d[6], a.l[2]);
        ZKW_SQR_REDUCE(nA, nB)
        ZKW_SQR_NEXT()
        // row 3: odd chain a3*a3, d5, d7; even chain d4, d6
        shift_mad_pairs_s1(nA[0], nB, B[0], A, a.l[3], d[5], d[7], a.l[3]);
        ZKW_SQR_COPY_B()
        mad_pairs_s2(nA, e[4], d[6], a.l[3]);
        ZKW_SQR_REDUCE(nA, nB)
        ZKW_SQR_NEXT()
        // row 4: odd chain d5, d7; even chain a4*a4, d6
        shift_mad_pairs_s2(nA[0], nB, B[0], A, e[5], d[7], a.l[4]);
        ZKW_SQR_COPY_B()
        mad_pairs_s2(nA, a.l[4], d[6], a.l[4]);
        ZKW_SQR_REDUCE(nA, nB)
        ZKW_SQR_NEXT()
        // row 5: odd chain a5*a5, d7; even chain d6
        shift_mad_pairs_s2(nA[0], nB, B[0], A, a.l[5], d[7], a.l[5]);
        ZKW_SQR_COPY_B()
        mad_pairs_s3(nA, e[6], a.l[5]);
        ZKW_SQR_REDUCE(nA, nB)
        ZKW_SQR_NEXT()
        // row 6: odd chain d7; even chain a6*a6
        shift_mad_pairs_s3(nA[0], nB, B[0], A, e[7], a.l[6]);
        ZKW_SQR_COPY_B()
        mad_pairs_s3(nA, a.l[6], a.l[6]);
        ZKW_SQR_REDUCE(nA, nB)
        ZKW_SQR_NEXT()
        // row 7: odd chain a7*a7; even chain empty
        shift_mad_pairs_s3(nA[0], nB, B[0], A, a.l[7], a.l[7]);
        ZKW_SQR_COPY_B()
        ZKW_SQR_REDUCE(nA, nB)
        ZKW_SQR_NEXT()
#undef ZKW_SQR_REDUCE
#undef ZKW_SQR_NEXT
#undef ZKW_SQR_COPY_B
        Fp r;
        asm("add.cc.u32 %0, %8, %16;\n\t"
            "addc.cc.u32 %1, %9, %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32 %7, %15, %23;"
            : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7])
            : "r"(A[1]), "r"(A[2]), "r"(A[3]), "r"(A[4]), "r"(A[5]), "r"(A[6]), "r"(A[7]), "r"(A[8]),
              "r"(B[0]), "r"(B[1]), "r"(B[2]), "r"(B[3]), "r"(B[4]), "r"(B[5]), "r"(B[6]), "r"(B[7]));
        return r;
#endif
    }
    // sum_k a_k * b_k * R^-1 mod m with ONE interleaved reduction for the K products: 64 (K + 1) wide multiplies instead of
    // 128 K.  T = sum a_k b_k, so the result (T + M m) / R < T / R + m; with m / R < 0.19 for both BN254 moduli:
    //   K = 2, operands <= 2m:  < 2.52 m  (below 4m < 2^256: callers bring it back with reduced_2m())
    //   K <= 5, operands <  m:  < 1.95 m  (the usual lazy range: normalized() makes it canonical)
    // The nine-limb accumulators take K + 1 row terms (< 2^256 each, the sum divided by 2^32 per row): far below 2^288.
    template <int K>
    __device__ __forceinline__ static Fp dot_lazy(const Fp (&a)[K], const Fp (&b)[K]) {
#ifndef __CUDA_ARCH__
        Fp r = zero();
        for (int k = 0; k < K; k++) r = r + a[k].normalized() * b[k].normalized();
        return r;
#else
        uint32_t A[9], B[9];
        mul_pairs(A, a[0].l[0], a[0].l[2], a[0].l[4], a[0].l[6], b[0].l[0]);
        mul_pairs(B, a[0].l[1], a[0].l[3], a[0].l[5], a[0].l[7], b[0].l[0]);
#pragma unroll
        for (int k = 1; k < K; k++) {
            mad_pairs(A, a[k].l[0], a[k].l[2], a[k].l[4], a[k].l[6], b[k].l[0]);
            mad_pairs(B, a[k].l[1], a[k].l[3], a[k].l[5], a[k].l[7], b[k].l[0]);
        }
        uint32_t m = A[0] * P::INV;
        mad_pairs(A, P::mod(0), P::mod(2), P::mod(4), P::mod(6), m);
        mad_pairs(B, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
#pragma unroll
        for (int i = 1; i < 8; i++) {
            uint32_t nA[9], nB[9];
            shift_mad_pairs(nA[0], nB, B[0], A, a[0].l[1], a[0].l[3], a[0].l[5], a[0].l[7], b[0].l[i]);
#pragma unroll
            for (int k = 1; k < 9; k++) nA[k] = B[k];
            mad_pairs(nA, a[0].l[0], a[0].l[2], a[0].l[4], a[0].l[6], b[0].l[i]);
#pragma unroll
            for (int k = 1; k < K; k++) {
                mad_pairs(nA, a[k].l[0], a[k].l[2], a[k].l[4], a[k].l[6], b[k].l[i]);
                mad_pairs(nB, a[k].l[1], a[k].l[3], a[k].l[5], a[k].l[7], b[k].l[i]);
            }
            m = nA[0] * P::INV;
            mad_pairs(nA, P::mod(0), P::mod(2), P::mod(4), P::mod(6), m);
            mad_pairs(nB, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
#pragma unroll
            for (int k = 0; k < 9; k++) { A[k] = nA[k]; B[k] = nB[k]; }
        }
        Fp r;
        asm("add.cc.u32 %0, %8, %16;\n\t"
            "addc.cc.u32 %1, %9, %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32 %7, %15, %23;"
            : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7])
            : "r"(A[1]), "r"(A[2]), "r"(A[3]), "r"(A[4]), "r"(A[5]), "r"(A[6]), "r"(A[7]), "r"(A[8]),
              "r"(B[0]), "r"(B[1]), "r"(B[2]), "r"(B[3]), "r"(B[4]), "r"(B[5]), "r"(B[6]), "r"(B[7]));
        return r;
#endif
    }
    // a b + c d for operands <= 2m: result < 2.52 m, see dot_lazy
    __device__ __forceinline__ static Fp mul2_lazy(const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
        const Fp x[2] = {a, c}, y[2] = {b, d};
        return dot_lazy<2>(x, y);
    }
    // 2m - a for a < 2m: -a in (0, 2m], one subtraction chain, no condition
    __device__ __forceinline__ Fp neg_2m() const {
#ifndef __CUDA_ARCH__
        return normalized().neg();
#else
        Fp r;
        asm("sub.cc.u32 %0, %8, %16;\n\t"
            "subc.cc.u32 %1, %9, %17;\n\t"
            "subc.cc.u32 %2, %10, %18;\n\t"
            "subc.cc.u32 %3, %11, %19;\n\t"
            "subc.cc.u32 %4, %12, %20;\n\t"
            "subc.cc.u32 %5, %13, %21;\n\t"
            "subc.cc.u32 %6, %14, %22;\n\t"
            "subc.u32 %7, %15, %23;"
            : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7])
            : "r"(mod2(0)), "r"(mod2(1)), "r"(mod2(2)), "r"(mod2(3)), "r"(mod2(4)), "r"(mod2(5)), "r"(mod2(6)), "r"(mod2(7)),
              "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]), "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7]));
        return r;
#endif
    }
    // a + b for a, b < 2m; result < 2m
    __device__ __forceinline__ static Fp add_lazy(const Fp& a, const Fp& b) {
#ifndef __CUDA_ARCH__
        return a.normalized() + b.normalized();
#else
        Fp r;
        asm("add.cc.u32 %0, %8, %16;\n\t"
            "addc.cc.u32 %1, %9, %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32 %7, %15, %23;"
            : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7])
            : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
              "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
        uint32_t t[8], borrow;   // r < 4m < 2^256: subtract 2m if r >= 2m
        asm("sub.cc.u32 %0, %9, %17;\n\t"
            "subc.cc.u32 %1, %10, %18;\n\t"
            "subc.cc.u32 %2, %11, %19;\n\t"
            "subc.cc.u32 %3, %12, %20;\n\t"
            "subc.cc.u32 %4, %13, %21;\n\t"
            "subc.cc.u32 %5, %14, %22;\n\t"
            "subc.cc.u32 %6, %15, %23;\n\t"
            "subc.cc.u32 %7, %16, %24;\n\t"
            "subc.u32 %8, 0, 0;"
            : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(borrow)
            : "r"(r.l[0]), "r"(r.l[1]), "r"(r.l[2]), "r"(r.l[3]), "r"(r.l[4]), "r"(r.l[5]), "r"(r.l[6]), "r"(r.l[7]),
              "r"(mod2(0)), "r"(mod2(1)), "r"(mod2(2)), "r"(mod2(3)), "r"(mod2(4)), "r"(mod2(5)), "r"(mod2(6)), "r"(mod2(7)));
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = borrow ? r.l[i] : t[i];
        return r;
#endif
    }
    // a - b for a, b < 2m; result < 2m
    __device__ __forceinline__ static Fp sub_lazy(const Fp& a, const Fp& b) {
#ifndef __CUDA_ARCH__
        return a.normalized() - b.normalized();
#else
        Fp r;
        uint32_t borrow;
        asm("sub.cc.u32 %0, %9, %17;\n\t"
            "subc.cc.u32 %1, %10, %18;\n\t"
            "subc.cc.u32 %2, %11, %19;\n\t"
            "subc.cc.u32 %3, %12, %20;\n\t"
            "subc.cc.u32 %4, %13, %21;\n\t"
            "subc.cc.u32 %5, %14, %22;\n\t"
            "subc.cc.u32 %6, %15, %23;\n\t"
            "subc.cc.u32 %7, %16, %24;\n\t"
            "subc.u32 %8, 0, 0;"
            : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7]), "=r"(borrow)
            : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
              "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
        asm("add.cc.u32 %0, %0, %8;\n\t"
            "addc.cc.u32 %1, %1, %9;\n\t"
            "addc.cc.u32 %2, %2, %10;\n\t"
            "addc.cc.u32 %3, %3, %11;\n\t"
            "addc.cc.u32 %4, %4, %12;\n\t"
            "addc.cc.u32 %5, %5, %13;\n\t"
            "addc.cc.u32 %6, %6, %14;\n\t"
            "addc.u32 %7, %7, %15;"
            : "+r"(r.l[0]), "+r"(r.l[1]), "+r"(r.l[2]), "+r"(r.l[3]), "+r"(r.l[4]), "+r"(r.l[5]), "+r"(r.l[6]), "+r"(r.l[7])
            : "r"(mod2(0) & borrow), "r"(mod2(1) & borrow), "r"(mod2(2) & borrow), "r"(mod2(3) & borrow),
              "r"(mod2(4) & borrow), "r"(mod2(5) & borrow), "r"(mod2(6) & borrow), "r"(mod2(7) & borrow));
        return r;
#endif
    }

    // ---- Harvey-style butterfly arithmetic: values in [0, 4m) (4m < 2^256 for both BN254 moduli) -------------------------
    // The NTT keeps its elements in [0, 4m) between stages: per butterfly ONE conditional subtraction (of 2m, on the
    // untwiddled input) instead of three (product, sum, difference).  mul_lazy accepts a < 4m, b < m: a*b/R + m < 2m.
    // [0, 4m) -> [0, 2m)
    __device__ __forceinline__ Fp reduced_2m() const {
#ifndef __CUDA_ARCH__
        return *this;
#else
        uint32_t t[8], borrow;
        asm("sub.cc.u32 %0, %9, %17;\n\t"
            "subc.cc.u32 %1, %10, %18;\n\t"
            "subc.cc.u32 %2, %11, %19;\n\t"
            "subc.cc.u32 %3, %12, %20;\n\t"
            "subc.cc.u32 %4, %13, %21;\n\t"
            "subc.cc.u32 %5, %14, %22;\n\t"
            "subc.cc.u32 %6, %15, %23;\n\t"
            "subc.cc.u32 %7, %16, %24;\n\t"
            "subc.u32 %8, 0, 0;"
            : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(borrow)
            : "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]), "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7]),
              "r"(mod2(0)), "r"(mod2(1)), "r"(mod2(2)), "r"(mod2(3)), "r"(mod2(4)), "r"(mod2(5)), "r"(mod2(6)), "r"(mod2(7)));
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = borrow ? l[i] : t[i];
        return r;
#endif
    }
    // a + b with no reduction (the caller guarantees a + b < 2^256)
    __device__ __forceinline__ static Fp add_raw(const Fp& a, const Fp& b) {
        Fp r;
#ifdef __CUDA_ARCH__
        asm("add.cc.u32 %0, %8, %16;\n\t"
            "addc.cc.u32 %1, %9, %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32 %7, %15, %23;"
            : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7])
            : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
              "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
#else
        r = a + b;
#endif
        return r;
    }
    // a - b + 2m for a, b < 2m: in (0, 4m), no condition
    __device__ __forceinline__ static Fp sub_plus_2m(const Fp& a, const Fp& b) {
        Fp r;
#ifdef __CUDA_ARCH__
        asm("add.cc.u32 %0, %8, %16;\n\t"
            "addc.cc.u32 %1, %9, %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32 %7, %15, %23;"
            : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7])
            : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
              "r"(mod2(0)), "r"(mod2(1)), "r"(mod2(2)), "r"(mod2(3)), "r"(mod2(4)), "r"(mod2(5)), "r"(mod2(6)), "r"(mod2(7)));
        asm("sub.cc.u32 %0, %0, %8;\n\t"
            "subc.cc.u32 %1, %1, %9;\n\t"
            "subc.cc.u32 %2, %2, %10;\n\t"
            "subc.cc.u32 %3, %3, %11;\n\t"
            "subc.cc.u32 %4, %4, %12;\n\t"
            "subc.cc.u32 %5, %5, %13;\n\t"
            "subc.cc.u32 %6, %6, %14;\n\t"
            "subc.u32 %7, %7, %15;"
            : "+r"(r.l[0]), "+r"(r.l[1]), "+r"(r.l[2]), "+r"(r.l[3]), "+r"(r.l[4]), "+r"(r.l[5]), "+r"(r.l[6]), "+r"(r.l[7])
            : "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
#else
        r = a - b;
#endif
        return r;
    }
    // [0, 4m) -> [0, m)
    __device__ __forceinline__ Fp normalized_4m() const {
        Fp r = reduced_2m();
        r.reduce_once();
        return r;
    }

    __host__ __device__ __forceinline__ Fp to_mont() const { return *this * r2(); }
    __host__ __device__ __forceinline__ Fp from_mont() const {
        Fp o = zero();
        o.l[0] = 1;
        return *this * o;
    }

    // a^e for a 64-bit exponent (square-and-multiply, MSB first)
    __host__ __device__ Fp pow(uint64_t e) const {
        Fp acc = one();
        int top = 63;
        while (top >= 0 && !((e >> top) & 1)) top--;  // skip leading zeros
        for (int i = top; i >= 0; i--) {
            acc = acc.sqr();
            if ((e >> i) & 1) acc = acc * *this;
        }
        return acc;
    }
    // Fermat inverse a^(m-2); inverse of zero is zero.
    __host__ __device__ Fp inv() const {
        Fp acc = one();
        for (int i = 7; i >= 0; i--) {
            uint32_t w = P::mod(i) - (i == 0 ? 2u : 0u);
            for (int b = 31; b >= 0; b--) {
                acc = acc.sqr();
                if ((w >> b) & 1) acc = acc * *this;
            }
        }
        return acc;
    }
};

using Fr = Fp<FrParams>;
using Fq = Fp<FqParams>;

}  // namespace zkw
