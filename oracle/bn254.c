/*
 * bn254.c — exported field / G1 shims of the CPU oracle, plus dev-SRS generation.
 * TEST INFRASTRUCTURE ONLY (see zkw_oracle.h).
 */
#include <omp.h>
#include <stdlib.h>
#include "bn254_internal.h"
#include "zkw_oracle.h"

int zko_max_threads(void) { return omp_get_max_threads(); }

void zko_fr_mul(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) { fr_mul(r, a, b); }
void zko_fr_add(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) { fr_add(r, a, b); }
void zko_fr_sub(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) { fr_sub(r, a, b); }
void zko_fr_inv(uint64_t r[4], const uint64_t a[4]) { fr_inv(r, a); }
void zko_fr_to_mont(uint64_t r[4], const uint64_t a[4]) { fr_to_mont(r, a); }
void zko_fr_from_mont(uint64_t r[4], const uint64_t a[4]) { fr_from_mont(r, a); }
void zko_fq_mul(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) { fq_mul(r, a, b); }
void zko_fq_add(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) { fq_add(r, a, b); }
void zko_fq_sub(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) { fq_sub(r, a, b); }
void zko_fq_inv(uint64_t r[4], const uint64_t a[4]) { fq_inv(r, a); }
void zko_fq_to_mont(uint64_t r[4], const uint64_t a[4]) { fq_to_mont(r, a); }
void zko_fq_from_mont(uint64_t r[4], const uint64_t a[4]) { fq_from_mont(r, a); }

void zko_fr_vec_to_mont(uint64_t* r, const uint64_t* a, size_t n) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) fr_to_mont(r + 4 * i, a + 4 * i);
}
void zko_fr_vec_from_mont(uint64_t* r, const uint64_t* a, size_t n) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) fr_from_mont(r + 4 * i, a + 4 * i);
}

static inline uint64_t splitmix64(uint64_t* s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

/* canonical value = (254 random bits) mod r, then to Montgomery form: matches pyref.SplitMix64.field */
void zko_fr_random(uint64_t* out, size_t n, uint64_t seed) {
    uint64_t s = seed;
    for (size_t i = 0; i < n; i++) {
        uint64_t v[4];
        for (int j = 0; j < 4; j++) v[j] = splitmix64(&s);
        v[3] &= 0x3FFFFFFFFFFFFFFFULL;
        /* v < 2^254 < 2r?  r ~ 0.756*2^254, so at most one subtraction */
        fr_cond_sub(v, v, 0);
        fr_to_mont(out + 4 * i, v);
    }
}

void zko_g1_generator(uint64_t out_xy[8]) {
    uint64_t one[4] = {1, 0, 0, 0}, two[4] = {2, 0, 0, 0};
    fq_to_mont(out_xy, one);
    fq_to_mont(out_xy + 4, two);
}

int zko_g1_is_on_curve(const uint64_t xy[8]) {
    const g1a_t* a = (const g1a_t*)xy;
    if (g1a_is_identity(a)) return 1;
    uint64_t y2[4], x3[4], three[4] = {3, 0, 0, 0};
    fq_to_mont(three, three);
    fq_sqr(y2, a->y);
    fq_sqr(x3, a->x); fq_mul(x3, x3, a->x); fq_add(x3, x3, three);
    return fq_eq(y2, x3);
}

void zko_g1_add(uint64_t r[12], const uint64_t a[12], const uint64_t b[12]) {
    g1_t t; g1_add(&t, (const g1_t*)a, (const g1_t*)b); memcpy(r, &t, 96);
}
void zko_g1_add_mixed(uint64_t r[12], const uint64_t a[12], const uint64_t b[8]) {
    g1_t t; g1_add_mixed(&t, (const g1_t*)a, (const g1a_t*)b); memcpy(r, &t, 96);
}
void zko_g1_double(uint64_t r[12], const uint64_t a[12]) {
    g1_t t; g1_double(&t, (const g1_t*)a); memcpy(r, &t, 96);
}
void zko_g1_mul(uint64_t r[12], const uint64_t a_xy[8], const uint64_t scalar[4]) {
    uint64_t s[4];
    fr_from_mont(s, scalar);
    g1_t acc; g1_set_identity(&acc);
    for (int i = 255; i >= 0; i--) {
        g1_double(&acc, &acc);
        if ((s[i >> 6] >> (i & 63)) & 1) g1_add_mixed(&acc, &acc, (const g1a_t*)a_xy);
    }
    memcpy(r, &acc, 96);
}
void zko_g1_to_affine(uint64_t out_xy[8], const uint64_t xyz[12]) {
    g1a_t t; g1_to_affine(&t, (const g1_t*)xyz); memcpy(out_xy, &t, 64);
}

/* C::Curve::batch_normalize: one shared inversion over the non-identity Z's */
void zko_g1_batch_to_affine(uint64_t* out_xy, const uint64_t* xyz, size_t m) {
    if (m == 0) return;
    uint64_t* z = (uint64_t*)malloc(m * 32);
    uint64_t* scratch = (uint64_t*)malloc(m * 32);
    for (size_t i = 0; i < m; i++) memcpy(z + 4 * i, xyz + 12 * i + 8, 32);
    fq_batch_inv(z, m, scratch);
    for (size_t i = 0; i < m; i++) {
        g1a_t* o = (g1a_t*)(out_xy + 8 * i);
        const g1_t* p = (const g1_t*)(xyz + 12 * i);
        if (g1_is_identity(p)) { fq_zero(o->x); fq_zero(o->y); continue; }
        uint64_t zi2[4], zi3[4];
        fq_sqr(zi2, z + 4 * i);
        fq_mul(zi3, zi2, z + 4 * i);
        fq_mul(o->x, p->x, zi2);
        fq_mul(o->y, p->y, zi3);
    }
    free(z);
    free(scratch);
}

/* Fixed-base table of the generator: T[w][d] = d * 2^(8w) * G, w < 32, d < 256 (affine). */
static g1a_t* g_fixed_table = NULL;
static void build_fixed_table(void) {
    if (g_fixed_table) return;
    g1_t* jac = (g1_t*)malloc(sizeof(g1_t) * 32 * 256);
    g1a_t g; zko_g1_generator((uint64_t*)&g);
    g1_t base; g1_from_affine(&base, &g);
    for (int w = 0; w < 32; w++) {
        g1_t* row = jac + w * 256;
        g1_set_identity(&row[0]);
        for (int d = 1; d < 256; d++) g1_add(&row[d], &row[d - 1], &base);
        for (int i = 0; i < 8; i++) g1_double(&base, &base);
    }
    g1a_t* t = (g1a_t*)malloc(sizeof(g1a_t) * 32 * 256);
    zko_g1_batch_to_affine((uint64_t*)t, (const uint64_t*)jac, 32 * 256);
    free(jac);
    g_fixed_table = t;
}

void zko_g1_fixed_base_mul(uint64_t* out_xy, const uint64_t* scalars, size_t n, int threads) {
#pragma omp critical(zko_fixed_table)
    build_fixed_table();
    if (threads <= 0) threads = omp_get_max_threads();
    g1_t* jac = (g1_t*)malloc(sizeof(g1_t) * n);
#pragma omp parallel for schedule(static) num_threads(threads)
    for (size_t i = 0; i < n; i++) {
        uint64_t s[4];
        fr_from_mont(s, scalars + 4 * i);
        g1_t acc; g1_set_identity(&acc);
        for (int w = 0; w < 32; w++) {
            unsigned d = (unsigned)((s[w >> 3] >> ((w & 7) * 8)) & 0xFF);
            if (d) g1_add_mixed(&acc, &acc, &g_fixed_table[w * 256 + d]);
        }
        jac[i] = acc;
    }
    zko_g1_batch_to_affine(out_xy, (const uint64_t*)jac, n);
    free(jac);
}

/* ParamsKZG::setup's monomial basis g[i] = tau^i * G (upstream draws tau from the rng; here it is
 * a caller-chosen dev value so that the SRS is reproducible on both sides of every test). */
void zko_srs_powers(uint64_t* out_xy, size_t n, const uint64_t tau[4], int threads) {
    uint64_t* pw = (uint64_t*)malloc(n * 32);
    uint64_t acc[4];
    fr_one(acc);
    for (size_t i = 0; i < n; i++) {
        fr_set(pw + 4 * i, acc);
        fr_mul(acc, acc, tau);
    }
    zko_g1_fixed_base_mul(out_xy, pw, n, threads);
    free(pw);
}
