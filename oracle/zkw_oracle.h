/*
 * zkw_oracle.h — CPU oracle for the Halo2 prover hot path (BN254 MSM / NTT / quotient).
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may link or load this library, and only as the checker / the timed
 * CPU arm.  The product (webauthn-halo2_b200/) never calls into it and has no CPU fallback.
 *
 * PARITY STATUS: **unpinned** at the MSM / NTT / quotient boundary.  The reference
 * (/root/reference) carries no known-answer for these functions: they live in un-vendored,
 * un-pinned crates (halo2-circuits/Cargo.toml:12-15 -> zkwebauthn/halo2-lib@main -> PSE
 * halo2_proofs tag v2023_01_20, halo2curves 0.3; zkwebauthn/snark-verifier@v2023_01_20_secp256r1),
 * and there is no Rust toolchain here to build them.  This library restates their published
 * algorithms; it is pinned by (a) the field/domain constants embedded in the reference's
 * generated verifier (proving-server/P256Verifier.yul:17-18,306-323,465-509), (b) an independent
 * Python big-integer restatement (oracle/pyref.py), and (c) end-to-end: proofs built from these
 * functions are accepted by a restated verifier that also accepts the reference's golden proof
 * (contracts/test/P256Account.t.sol:120) under the reference's own Yul.
 *
 * All field elements: [u64;4] little-endian limbs, Montgomery form (halo2curves layout).
 */
#ifndef ZKW_ORACLE_H
#define ZKW_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#include "../include/zkw_b200.h" /* POD structs only (zkw_circuit_shape, zkw_quotient_inputs) */

#ifdef __cplusplus
extern "C" {
#endif

/* ---- scalar-field / base-field element ops (for tests and host-side glue of the tests) ---- */
void zko_fr_mul(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]);
void zko_fr_add(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]);
void zko_fr_sub(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]);
void zko_fr_inv(uint64_t r[4], const uint64_t a[4]);
void zko_fr_to_mont(uint64_t r[4], const uint64_t a[4]);
void zko_fr_from_mont(uint64_t r[4], const uint64_t a[4]);
void zko_fq_mul(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]);
void zko_fq_add(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]);
void zko_fq_sub(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]);
void zko_fq_inv(uint64_t r[4], const uint64_t a[4]);
void zko_fq_to_mont(uint64_t r[4], const uint64_t a[4]);
void zko_fq_from_mont(uint64_t r[4], const uint64_t a[4]);
/* vector forms: n elements, in place where r == a */
void zko_fr_vec_to_mont(uint64_t* r, const uint64_t* a, size_t n);
void zko_fr_vec_from_mont(uint64_t* r, const uint64_t* a, size_t n);
/* deterministic uniform-ish Fr / scalars in Montgomery form from a splitmix64 stream
 * (same stream as oracle/pyref.py::SplitMix64.field) */
void zko_fr_random(uint64_t* out, size_t n, uint64_t seed);

/* ---- G1 (y^2 = x^3 + 3 over Fq) ----------------------------------------------------------- */
void zko_g1_generator(uint64_t out_xy[8]);
int zko_g1_is_on_curve(const uint64_t xy[8]);
void zko_g1_add(uint64_t r_xyz[12], const uint64_t a_xyz[12], const uint64_t b_xyz[12]);
void zko_g1_add_mixed(uint64_t r_xyz[12], const uint64_t a_xyz[12], const uint64_t b_xy[8]);
void zko_g1_double(uint64_t r_xyz[12], const uint64_t a_xyz[12]);
/* scalar in Montgomery form (as every Fr at this boundary) */
void zko_g1_mul(uint64_t r_xyz[12], const uint64_t a_xy[8], const uint64_t scalar[4]);
void zko_g1_to_affine(uint64_t out_xy[8], const uint64_t xyz[12]);
void zko_g1_batch_to_affine(uint64_t* out_xy, const uint64_t* xyz, size_t m);
/* bases[i] = (tau^i) * G for i < n, affine Montgomery: ParamsKZG::setup's g (tau in Montgomery form) */
void zko_srs_powers(uint64_t* out_xy, size_t n, const uint64_t tau[4], int threads);
/* out[i] = scalars[i] * G (fixed-base, windowed), affine */
void zko_g1_fixed_base_mul(uint64_t* out_xy, const uint64_t* scalars, size_t n, int threads);

/* ---- MSM: halo2_proofs::arithmetic::best_multiexp ---------------------------------------- */
/* `threads` plays the role of rayon's current_num_threads(): one contiguous chunk per thread,
 * each through multiexp_serial.  threads <= 0 means omp_get_max_threads(). */
void zko_best_multiexp(uint64_t out_xyz[12], const uint64_t* scalars, const uint64_t* bases_xy, size_t n, int threads);
void zko_msm_naive(uint64_t out_xyz[12], const uint64_t* scalars, const uint64_t* bases_xy, size_t n);

/* ---- NTT: halo2_proofs::arithmetic::best_fft --------------------------------------------- */
void zko_best_fft(uint64_t* a, unsigned log_n, const uint64_t omega[4], int threads);

/* ---- poly::EvaluationDomain ---------------------------------------------------------------- */
typedef struct {
    unsigned k, ext_k, quotient_poly_degree;
    uint64_t omega[4], omega_inv[4], ext_omega[4], ext_omega_inv[4];
    uint64_t g_coset[4], g_coset_inv[4], ifft_divisor[4], ext_ifft_divisor[4];
    uint64_t t_evaluations[16][4]; /* 2^(ext_k-k) <= 16 entries */
} zko_domain;
void zko_domain_new(zko_domain* d, unsigned cs_degree, unsigned k);
void zko_lagrange_to_coeff(const zko_domain* d, uint64_t* a, int threads);
void zko_coeff_to_lagrange(const zko_domain* d, uint64_t* a, int threads);
void zko_coeff_to_extended(const zko_domain* d, const uint64_t* coeffs, uint64_t* out, int threads);
void zko_extended_to_coeff(const zko_domain* d, uint64_t* a, int threads);
void zko_divide_by_vanishing_poly(const zko_domain* d, uint64_t* a, int threads);

/* ---- Evaluator::evaluate_h (+ divide_by_vanishing_poly) for the ECDSA circuit shape ---------- */
int zko_quotient_ecdsa(const zkw_quotient_inputs* in, uint64_t* h_ext, int threads);

int zko_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
