/*
 * ntt.c — CPU oracle restatement of halo2_proofs::arithmetic::best_fft and the
 * poly::EvaluationDomain transforms built on it.  TEST INFRASTRUCTURE ONLY (see zkw_oracle.h).
 *
 * Reached by the reference only through create_proof / keygen
 * (halo2-circuits/src/ecc/ecdsa_p256.rs:259-260, 366-373, 416-423, 555-562); the crate is an
 * un-vendored dependency (Cargo.toml:12-13), so the contracts are restated:
 *   best_fft(a, omega, log_n): in place, natural order in/out, a[i] <- sum_j a[j] omega^(ij);
 *       bit-reversal permutation, then log_n radix-2 decimation-in-time stages using the
 *       precomputed table twiddles[i] = omega^i, i < n/2.  (Upstream splits the butterflies
 *       recursively across rayon threads; the arithmetic and the result are identical.)
 *   EvaluationDomain::new(j, k): quotient_poly_degree = j - 1; extended_k minimal with
 *       2^extended_k >= n (j-1); extended_omega = ROOT_OF_UNITY^(2^(S - extended_k));
 *       omega = extended_omega^(2^(extended_k - k)); g_coset = ZETA, g_coset_inv = ZETA^2;
 *       t_evaluations[i] = 1 / ((ZETA * extended_omega^i)^n - 1), i < 2^(extended_k - k).
 *   lagrange_to_coeff = best_fft(omega^-1) * n^-1;
 *   coeff_to_extended = a_i *= ZETA^(i mod 3); zero-extend; best_fft(extended_omega);
 *   extended_to_coeff = best_fft(extended_omega^-1) * 2^-extended_k; a_i *= ZETA^-(i mod 3);
 *   divide_by_vanishing_poly: a_i *= t_evaluations[i mod len].
 * Pinned constants: n^-1 and omega^-j for k = 17 (P256Verifier.yul:307-323), tests/test_oracle_constants.py.
 */
#include <omp.h>
#include <stdlib.h>
#include "bn254_internal.h"
#include "zkw_oracle.h"

static inline size_t bitrev(size_t x, unsigned l) {
    size_t r = 0;
    for (unsigned i = 0; i < l; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

void zko_best_fft(uint64_t* a, unsigned log_n, const uint64_t omega[4], int threads) {
    if (threads <= 0) threads = omp_get_max_threads();
    size_t n = (size_t)1 << log_n;
    if (log_n == 0) return;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (size_t k = 0; k < n; k++) {
        size_t rk = bitrev(k, log_n);
        if (k < rk) {
            uint64_t t[4];
            fr_set(t, a + 4 * k); fr_set(a + 4 * k, a + 4 * rk); fr_set(a + 4 * rk, t);
        }
    }
    /* twiddles[i] = omega^i, i < n/2: blocks of 1024 seeded by a direct power, filled serially */
    size_t half_n = n / 2;
    uint64_t* tw = (uint64_t*)malloc((half_n ? half_n : 1) * 32);
    const size_t blk = 1024;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (size_t b0 = 0; b0 < half_n; b0 += blk) {
        uint64_t w[4];
        fr_pow_u64(w, omega, (uint64_t)b0);
        size_t hi = b0 + blk > half_n ? half_n : b0 + blk;
        for (size_t i = b0; i < hi; i++) { fr_set(tw + 4 * i, w); fr_mul(w, w, omega); }
    }
    size_t chunk = 2, tchunk = half_n;
    for (unsigned s = 0; s < log_n; s++) {
        size_t half = chunk / 2;
#pragma omp parallel for schedule(static) num_threads(threads)
        for (size_t bf = 0; bf < half_n; bf++) {
            size_t grp = bf / half, i = bf % half;
            uint64_t* lo = a + 4 * (grp * chunk + i);
            uint64_t* hi = lo + 4 * half;
            uint64_t t[4], u[4];
            if (i == 0) fr_set(t, hi); else fr_mul(t, hi, tw + 4 * (i * tchunk));
            fr_set(u, lo);
            fr_add(lo, u, t);
            fr_sub(hi, u, t);
        }
        chunk *= 2;
        tchunk /= 2;
    }
    free(tw);
}

void zko_domain_new(zko_domain* d, unsigned cs_degree, unsigned k) {
    memset(d, 0, sizeof(*d));
    d->k = k;
    d->quotient_poly_degree = cs_degree - 1;
    unsigned ek = k;
    while (((uint64_t)1 << ek) < ((uint64_t)1 << k) * d->quotient_poly_degree) ek++;
    d->ext_k = ek;
    uint64_t w[4];
    fr_set(w, FR_ROOT_OF_UNITY_M);
    for (unsigned i = ek; i < FR_TWO_ADICITY; i++) fr_sqr(w, w);
    fr_set(d->ext_omega, w);
    fr_inv(d->ext_omega_inv, w);
    for (unsigned i = k; i < ek; i++) fr_sqr(w, w);
    fr_set(d->omega, w);
    fr_inv(d->omega_inv, w);
    fr_set(d->g_coset, FR_ZETA_M);
    fr_sqr(d->g_coset_inv, FR_ZETA_M);
    uint64_t nn[4] = {(uint64_t)1 << k, 0, 0, 0}, en[4] = {(uint64_t)1 << ek, 0, 0, 0};
    fr_to_mont(nn, nn); fr_inv(d->ifft_divisor, nn);
    fr_to_mont(en, en); fr_inv(d->ext_ifft_divisor, en);
    unsigned m = 1u << (ek - k);
    uint64_t cur[4], one[4];
    fr_set(cur, d->g_coset);
    fr_one(one);
    for (unsigned i = 0; i < m && i < 16; i++) {
        uint64_t t[4];
        fr_pow_u64(t, cur, (uint64_t)1 << k);
        fr_sub(t, t, one);
        fr_inv(d->t_evaluations[i], t);
        fr_mul(cur, cur, d->ext_omega);
    }
}

static void scale_all(uint64_t* a, size_t n, const uint64_t s[4], int threads) {
#pragma omp parallel for schedule(static) num_threads(threads)
    for (size_t i = 0; i < n; i++) fr_mul(a + 4 * i, a + 4 * i, s);
}

void zko_lagrange_to_coeff(const zko_domain* d, uint64_t* a, int threads) {
    if (threads <= 0) threads = omp_get_max_threads();
    zko_best_fft(a, d->k, d->omega_inv, threads);
    scale_all(a, (size_t)1 << d->k, d->ifft_divisor, threads);
}

void zko_coeff_to_lagrange(const zko_domain* d, uint64_t* a, int threads) {
    zko_best_fft(a, d->k, d->omega, threads);
}

void zko_coeff_to_extended(const zko_domain* d, const uint64_t* coeffs, uint64_t* out, int threads) {
    if (threads <= 0) threads = omp_get_max_threads();
    size_t n = (size_t)1 << d->k, en = (size_t)1 << d->ext_k;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (size_t i = 0; i < en; i++) {
        if (i >= n) { fr_zero(out + 4 * i); continue; }
        switch (i % 3) {
            case 0: fr_set(out + 4 * i, coeffs + 4 * i); break;
            case 1: fr_mul(out + 4 * i, coeffs + 4 * i, d->g_coset); break;
            default: fr_mul(out + 4 * i, coeffs + 4 * i, d->g_coset_inv); break;
        }
    }
    zko_best_fft(out, d->ext_k, d->ext_omega, threads);
}

void zko_extended_to_coeff(const zko_domain* d, uint64_t* a, int threads) {
    if (threads <= 0) threads = omp_get_max_threads();
    size_t en = (size_t)1 << d->ext_k;
    zko_best_fft(a, d->ext_k, d->ext_omega_inv, threads);
#pragma omp parallel for schedule(static) num_threads(threads)
    for (size_t i = 0; i < en; i++) {
        fr_mul(a + 4 * i, a + 4 * i, d->ext_ifft_divisor);
        switch (i % 3) {
            case 0: break;
            case 1: fr_mul(a + 4 * i, a + 4 * i, d->g_coset_inv); break;
            default: fr_mul(a + 4 * i, a + 4 * i, d->g_coset); break;
        }
    }
}

void zko_divide_by_vanishing_poly(const zko_domain* d, uint64_t* a, int threads) {
    if (threads <= 0) threads = omp_get_max_threads();
    size_t en = (size_t)1 << d->ext_k, m = (size_t)1 << (d->ext_k - d->k);
#pragma omp parallel for schedule(static) num_threads(threads)
    for (size_t i = 0; i < en; i++) fr_mul(a + 4 * i, a + 4 * i, d->t_evaluations[i % m]);
}
