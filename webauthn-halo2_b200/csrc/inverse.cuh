// inverse.cuh — modular inversion in Fp by an approximated binary GCD (Pornin, "Optimized Binary GCD for
// Modular Inversion", 2020), branch-free so that all 32 lanes of a warp run one instruction stream.
//
// Why it exists: the batched-affine bucket accumulation of msm.cu shares one inversion between the
// additions a thread performs.  Fermat's a^(p-2) costs ~290 Montgomery products on the multiplier pipe that
// bounds every kernel of this library; the binary GCD below costs ~80 wide multiplies per outer round
// (17 rounds, ~11 product equivalents in total) and otherwise runs on the integer add/logic pipe, which
// the products leave mostly idle.
//
// Round structure (k = 32): 64-bit approximations of (a, b) made of their top 33 and low 31 bits drive 30
// divsteps that only touch 64-bit words and produce a 2x2 matrix (f0 g0; f1 g1) with |f|+|g| <= 2^30; the
// matrix is then applied to the full-width (a, b) (exact division by 2^30) and to (u, v) modulo m
// (Montgomery-style division by 2^30), keeping a = u*y, b = v*y (mod m).  After ceil((2*254 - 1) / 30) = 17
// rounds a = 0 and b = gcd = 1, so v = y^-1.
#pragma once
#include "field.cuh"

namespace zkw {

namespace bingcd {

constexpr int kInner = 30;   // divsteps per round
constexpr int kRounds = 17;  // 17 * 30 = 510 >= 2 * 254 - 1

// out[0..8] = x[0..7] * f (f < 2^31)
__host__ __device__ __forceinline__ void mul_small(uint32_t out[9], const uint32_t x[8], uint32_t f) {
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)x[i] * f;
        out[i] = (uint32_t)c;
        c >>= 32;
    }
    out[8] = (uint32_t)c;
}

// x = (x ^ mask) - mask over 9 limbs: two's complement negation when mask = ~0, identity when mask = 0
__host__ __device__ __forceinline__ void cond_negate9(uint32_t x[9], uint32_t mask) {
    uint64_t c = mask & 1u;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        c += (uint64_t)(x[i] ^ mask);
        x[i] = (uint32_t)c;
        c >>= 32;
    }
}

// t = sf * |f| * x + sg * |g| * y as a 288-bit two's complement integer; returns the sign mask of t and
// leaves |t| >> 30 in out[0..7] (the caller guarantees exact divisibility and |t| < 2^286)
__host__ __device__ __forceinline__ uint32_t lincomb_shift(uint32_t out[8], const uint32_t x[8], const uint32_t y[8], int32_t f, int32_t g) {
    const uint32_t sf = (uint32_t)(f >> 31), sg = (uint32_t)(g >> 31);
    const uint32_t af = ((uint32_t)f ^ sf) - sf, ag = ((uint32_t)g ^ sg) - sg;
    uint32_t X[9], Y[9];
    mul_small(X, x, af);
    mul_small(Y, y, ag);
    cond_negate9(X, sf);
    cond_negate9(Y, sg);
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        c += (uint64_t)X[i] + Y[i];
        X[i] = (uint32_t)c;
        c >>= 32;
    }
    const uint32_t neg = (uint32_t)((int32_t)X[8] >> 31);
    cond_negate9(X, neg);
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = (X[i] >> kInner) | (X[i + 1] << (32 - kInner));
    return neg;
}

// out = (|f| * xs + |g| * ys) / 2^30 mod m, with xs = f < 0 ? m - x : x (so the signed combination is
// formed from non-negative terms), x, y < m, |f| + |g| <= 2^30; result < m
template <class P>
__host__ __device__ __forceinline__ void lincomb_mod(uint32_t out[8], const uint32_t x[8], const uint32_t y[8], int32_t f, int32_t g) {
    const uint32_t sf = (uint32_t)(f >> 31), sg = (uint32_t)(g >> 31);
    const uint32_t af = ((uint32_t)f ^ sf) - sf, ag = ((uint32_t)g ^ sg) - sg;
    uint32_t xs[8], ys[8];
    {
        // m - x and m - y, selected by the signs
        int64_t bx = 0, by = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            bx += (int64_t)P::mod(i) - (int64_t)x[i];
            by += (int64_t)P::mod(i) - (int64_t)y[i];
            xs[i] = (sf & (uint32_t)bx) | (~sf & x[i]);
            ys[i] = (sg & (uint32_t)by) | (~sg & y[i]);
            bx >>= 32;
            by >>= 32;
        }
    }
    uint32_t T[9];
    {
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            c += (uint64_t)xs[i] * af;
            const uint64_t d = (uint64_t)ys[i] * ag;
            // c + d can exceed 64 bits only if both are huge; split the addition to keep the carry
            const uint64_t s = c + d;
            const uint64_t carry = s < c ? 1 : 0;
            T[i] = (uint32_t)s;
            c = (s >> 32) | (carry << 32);
        }
        T[8] = (uint32_t)c;
    }
    // Montgomery step for 2^30: q = -T * m^-1 mod 2^30, T = (T + q * m) / 2^30
    const uint32_t q = (T[0] * P::INV) & ((1u << kInner) - 1u);
    {
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            c += (uint64_t)P::mod(i) * q + T[i];
            T[i] = (uint32_t)c;
            c >>= 32;
        }
        c += T[8];
        T[8] = (uint32_t)c;
    }
    uint32_t r[8];
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = (T[i] >> kInner) | (T[i + 1] << (32 - kInner));
    // r < 2m: one conditional subtraction
    uint32_t t[8];
    int64_t b = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        b += (int64_t)r[i] - (int64_t)P::mod(i);
        t[i] = (uint32_t)b;
        b >>= 32;
    }
    const uint32_t keep = (uint32_t)b;  // all ones if r < m
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = (keep & r[i]) | (~keep & t[i]);
}

}  // namespace bingcd

// Inverse of a Montgomery-form element, in Montgomery form; the inverse of zero is zero.
template <class P>
__host__ __device__ inline Fp<P> fp_inv_bingcd(const Fp<P>& y) {
    using namespace bingcd;
    uint32_t a[8], b[8], u[8], v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = y.l[i]; b[i] = P::mod(i); u[i] = 0; v[i] = 0; }
    u[0] = 1;
#pragma unroll 1
    for (int round = 0; round < kRounds; round++) {
        // ---- 64-bit approximations: top 33 bits (aligned on the longer of a, b) and low 31 bits ----
        uint32_t at = a[1], am = a[0], al = 0, bt = b[1], bm = b[0], bl = 0;
        uint32_t big = 0;  // all ones once a limb above index 1 is non-zero
#pragma unroll
        for (int i = 2; i < 8; i++) {
            const uint32_t nz = (a[i] | b[i]) != 0 ? 0xffffffffu : 0u;
            at = (nz & a[i]) | (~nz & at); am = (nz & a[i - 1]) | (~nz & am); al = (nz & a[i - 2]) | (~nz & al);
            bt = (nz & b[i]) | (~nz & bt); bm = (nz & b[i - 1]) | (~nz & bm); bl = (nz & b[i - 2]) | (~nz & bl);
            big |= nz;
        }
        uint32_t top = at | bt;
        int s = 0;
#ifdef __CUDA_ARCH__
        s = __clz((int)top);
#else
        s = top ? __builtin_clz(top) : 32;
#endif
        s &= 31;  // top == 0 only when big == 0, where the shift is not used
        // (t:m:l) << s, top 64 bits
        const uint32_t ah1 = s ? (at << s) | (am >> (32 - s)) : at, ah0 = s ? (am << s) | (al >> (32 - s)) : am;
        const uint32_t bh1 = s ? (bt << s) | (bm >> (32 - s)) : bt, bh0 = s ? (bm << s) | (bl >> (32 - s)) : bm;
        uint32_t xh = (big & ah1) | (~big & a[1]);
        uint32_t xl = (big & ((ah0 & 0x80000000u) | (a[0] & 0x7fffffffu))) | (~big & a[0]);
        uint32_t yh = (big & bh1) | (~big & b[1]);
        uint32_t yl = (big & ((bh0 & 0x80000000u) | (b[0] & 0x7fffffffu))) | (~big & b[0]);
        // ---- 30 divsteps on the approximations ----
        int32_t f0 = 1, g0 = 0, f1 = 0, g1 = 1;
#pragma unroll 2
        for (int i = 0; i < kInner; i++) {
            const uint32_t odd = 0u - (xl & 1u);
            // borrow of x - y
            const uint32_t lt = (xh < yh || (xh == yh && xl < yl)) ? 0xffffffffu : 0u;
            const uint32_t sw = odd & lt;
            uint32_t t;
            t = (xl ^ yl) & sw; xl ^= t; yl ^= t;
            t = (xh ^ yh) & sw; xh ^= t; yh ^= t;
            t = ((uint32_t)f0 ^ (uint32_t)f1) & sw; f0 = (int32_t)((uint32_t)f0 ^ t); f1 = (int32_t)((uint32_t)f1 ^ t);
            t = ((uint32_t)g0 ^ (uint32_t)g1) & sw; g0 = (int32_t)((uint32_t)g0 ^ t); g1 = (int32_t)((uint32_t)g1 ^ t);
            // x -= y & odd
            const uint32_t sl = yl & odd, sh = yh & odd;
            const uint32_t br = xl < sl ? 1u : 0u;
            xl -= sl;
            xh = xh - sh - br;
            f0 -= (int32_t)((uint32_t)f1 & odd);
            g0 -= (int32_t)((uint32_t)g1 & odd);
            // x >>= 1; second row doubles
            xl = (xl >> 1) | (xh << 31);
            xh >>= 1;
            f1 = (int32_t)((uint32_t)f1 << 1);
            g1 = (int32_t)((uint32_t)g1 << 1);
        }
        // ---- apply the matrix ----
        uint32_t na[8], nb[8];
        const uint32_t nega = lincomb_shift(na, a, b, f0, g0);
        const uint32_t negb = lincomb_shift(nb, a, b, f1, g1);
        f0 = (int32_t)(((uint32_t)f0 ^ nega) - nega); g0 = (int32_t)(((uint32_t)g0 ^ nega) - nega);
        f1 = (int32_t)(((uint32_t)f1 ^ negb) - negb); g1 = (int32_t)(((uint32_t)g1 ^ negb) - negb);
        uint32_t nu[8], nv[8];
        lincomb_mod<P>(nu, u, v, f0, g0);
        lincomb_mod<P>(nv, u, v, f1, g1);
#pragma unroll
        for (int i = 0; i < 8; i++) { a[i] = na[i]; b[i] = nb[i]; u[i] = nu[i]; v[i] = nv[i]; }
    }
    // b == 1 unless y == 0 (then b == m); v = Y^-1 as a plain residue.  Montgomery form of the inverse of
    // the Montgomery-form input: Y^-1 * R^2 = montmul(montmul(v, R^2), R^2).
    uint32_t ok = b[0] ^ 1u;
#pragma unroll
    for (int i = 1; i < 8; i++) ok |= b[i];
    const uint32_t mask = ok == 0 ? 0xffffffffu : 0u;
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = v[i] & mask;
    const Fp<P> r2 = Fp<P>::r2();
    return (r * r2) * r2;
}

}  // namespace zkw
