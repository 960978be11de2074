/*
 * bn254_internal.h — Fr / Fq instantiations and G1 Jacobian arithmetic for the CPU oracle.
 * TEST INFRASTRUCTURE ONLY (see zkw_oracle.h).  Restates halo2curves::bn256::{Fr,Fq,G1Affine,G1}
 * (imported by the reference at halo2-circuits/src/ecc/ecdsa_p256.rs:27); moduli are the
 * reference's f_q / f_p (proving-server/P256Verifier.yul:17-18).
 */
#ifndef ZKO_BN254_INTERNAL_H
#define ZKO_BN254_INTERNAL_H
#include <stdint.h>
#include <string.h>

/* ---- Fr: scalar field, r = 0x30644e72...f0000001 (yul:18 f_q) ---- */
#define FNAME(x) fr_##x
#define FMOD {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL}
#define FINV 0xc2e1f593efffffffULL
#define FR2 {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}
#define FONE {0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL}
#include "field_impl.h"
#undef FNAME
#undef FMOD
#undef FINV
#undef FR2
#undef FONE

/* ---- Fq: base field, p = 0x30644e72...d87cfd47 (yul:17 f_p) ---- */
#define FNAME(x) fq_##x
#define FMOD {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL}
#define FINV 0x87d20782e4866389ULL
#define FR2 {0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL}
#define FONE {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL}
#include "field_impl.h"
#undef FNAME
#undef FMOD
#undef FINV
#undef FR2
#undef FONE

/* Fr domain constants, Montgomery form: ROOT_OF_UNITY = 7^((r-1)/2^28), ZETA = 7^((r-1)/3),
 * DELTA = 7^(2^28)  (halo2curves bn256 Fr: MULTIPLICATIVE_GENERATOR = 7, S = 28; delta is pinned
 * by yul:465) */
static const uint64_t FR_ROOT_OF_UNITY_M[4] = {0x9632c7c5b639feb8ULL, 0x985ce3400d0ff299ULL, 0xb2dd880001b0ecd8ULL, 0x1d69070d6d98ce29ULL};
static const uint64_t FR_ZETA_M[4] = {0x93e7cede4a0329b3ULL, 0x7d4fdca77a96c167ULL, 0x8be4ba08b19a750aULL, 0x1cbd5653a5661c25ULL};
static const uint64_t FR_DELTA_M[4] = {0x9a0c322befd78855ULL, 0x46e82d14249b563cULL, 0x5983a663e0b0b7a7ULL, 0x22ab452baaa111adULL};
#define FR_TWO_ADICITY 28

/* ---- G1: Jacobian (X,Y,Z), x = X/Z^2, y = Y/Z^3; identity Z = 0; affine identity (0,0) ---- */
typedef struct { uint64_t x[4], y[4], z[4]; } g1_t;
typedef struct { uint64_t x[4], y[4]; } g1a_t;

static inline int g1_is_identity(const g1_t* a) { return fq_is_zero(a->z); }
static inline int g1a_is_identity(const g1a_t* a) { return fq_is_zero(a->x) && fq_is_zero(a->y); }
static inline void g1_set_identity(g1_t* r) { fq_zero(r->x); fq_one(r->y); fq_zero(r->z); }
static inline void g1_from_affine(g1_t* r, const g1a_t* a) {
    if (g1a_is_identity(a)) { g1_set_identity(r); return; }
    fq_set(r->x, a->x); fq_set(r->y, a->y); fq_one(r->z);
}

/* doubling, a = 0: "dbl-2009-l" */
static inline void g1_double(g1_t* r, const g1_t* p) {
    if (g1_is_identity(p)) { g1_set_identity(r); return; }
    uint64_t A[4], B[4], C[4], D[4], E[4], F[4], t[4], x3[4], y3[4], z3[4];
    fq_sqr(A, p->x);
    fq_sqr(B, p->y);
    fq_sqr(C, B);
    fq_add(t, p->x, B); fq_sqr(t, t); fq_sub(t, t, A); fq_sub(t, t, C); fq_dbl(D, t);
    fq_dbl(E, A); fq_add(E, E, A);
    fq_sqr(F, E);
    fq_dbl(t, D); fq_sub(x3, F, t);
    fq_mul(z3, p->y, p->z); fq_dbl(z3, z3);
    fq_sub(t, D, x3); fq_mul(y3, E, t);
    fq_dbl(t, C); fq_dbl(t, t); fq_dbl(t, t); fq_sub(y3, y3, t);
    fq_set(r->x, x3); fq_set(r->y, y3); fq_set(r->z, z3);
}

/* general addition: "add-2007-bl", falling back to doubling / identity on equal x */
static inline void g1_add(g1_t* r, const g1_t* p, const g1_t* q) {
    if (g1_is_identity(p)) { *r = *q; return; }
    if (g1_is_identity(q)) { *r = *p; return; }
    uint64_t z1z1[4], z2z2[4], u1[4], u2[4], s1[4], s2[4], h[4], i[4], j[4], rr[4], v[4], t[4];
    fq_sqr(z1z1, p->z); fq_sqr(z2z2, q->z);
    fq_mul(u1, p->x, z2z2); fq_mul(u2, q->x, z1z1);
    fq_mul(s1, p->y, q->z); fq_mul(s1, s1, z2z2);
    fq_mul(s2, q->y, p->z); fq_mul(s2, s2, z1z1);
    if (fq_eq(u1, u2)) {
        if (fq_eq(s1, s2)) { g1_double(r, p); } else { g1_set_identity(r); }
        return;
    }
    fq_sub(h, u2, u1);
    fq_dbl(i, h); fq_sqr(i, i);
    fq_mul(j, h, i);
    fq_sub(rr, s2, s1); fq_dbl(rr, rr);
    fq_mul(v, u1, i);
    uint64_t x3[4], y3[4], z3[4];
    fq_sqr(x3, rr); fq_sub(x3, x3, j); fq_dbl(t, v); fq_sub(x3, x3, t);
    fq_sub(t, v, x3); fq_mul(y3, rr, t); fq_mul(t, s1, j); fq_dbl(t, t); fq_sub(y3, y3, t);
    fq_add(z3, p->z, q->z); fq_sqr(z3, z3); fq_sub(z3, z3, z1z1); fq_sub(z3, z3, z2z2); fq_mul(z3, z3, h);
    fq_set(r->x, x3); fq_set(r->y, y3); fq_set(r->z, z3);
}

/* mixed addition: "madd-2007-bl" */
static inline void g1_add_mixed(g1_t* r, const g1_t* p, const g1a_t* q) {
    if (g1a_is_identity(q)) { *r = *p; return; }
    if (g1_is_identity(p)) { g1_from_affine(r, q); return; }
    uint64_t z1z1[4], u2[4], s2[4], h[4], hh[4], i[4], j[4], rr[4], v[4], t[4];
    fq_sqr(z1z1, p->z);
    fq_mul(u2, q->x, z1z1);
    fq_mul(s2, q->y, p->z); fq_mul(s2, s2, z1z1);
    if (fq_eq(p->x, u2)) {
        if (fq_eq(p->y, s2)) { g1_double(r, p); } else { g1_set_identity(r); }
        return;
    }
    fq_sub(h, u2, p->x);
    fq_sqr(hh, h);
    fq_dbl(i, hh); fq_dbl(i, i);
    fq_mul(j, h, i);
    fq_sub(rr, s2, p->y); fq_dbl(rr, rr);
    fq_mul(v, p->x, i);
    uint64_t x3[4], y3[4], z3[4];
    fq_sqr(x3, rr); fq_sub(x3, x3, j); fq_dbl(t, v); fq_sub(x3, x3, t);
    fq_sub(t, v, x3); fq_mul(y3, rr, t); fq_mul(t, p->y, j); fq_dbl(t, t); fq_sub(y3, y3, t);
    fq_add(z3, p->z, h); fq_sqr(z3, z3); fq_sub(z3, z3, z1z1); fq_sub(z3, z3, hh);
    fq_set(r->x, x3); fq_set(r->y, y3); fq_set(r->z, z3);
}

static inline void g1_to_affine(g1a_t* r, const g1_t* p) {
    if (g1_is_identity(p)) { fq_zero(r->x); fq_zero(r->y); return; }
    uint64_t zi[4], zi2[4], zi3[4];
    fq_inv(zi, p->z);
    fq_sqr(zi2, zi);
    fq_mul(zi3, zi2, zi);
    fq_mul(r->x, p->x, zi2);
    fq_mul(r->y, p->y, zi3);
}

#endif
