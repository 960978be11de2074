/*
 * field_impl.h — 4x64-bit Montgomery prime-field template for the CPU oracle.
 * TEST INFRASTRUCTURE ONLY (see oracle/README in zkw_oracle.h).
 *
 * Restates halo2curves' `field_arithmetic!` contract for bn256::{Fr,Fq} (the reference imports
 * them at halo2-circuits/src/ecc/ecdsa_p256.rs:27): elements are [u64;4] little-endian limbs in
 * Montgomery form with R = 2^256, always fully reduced (< modulus).
 *
 * Instantiate with:  #define FNAME(x) fr_##x   #define FMOD ...  #define FINV ...
 *                    #define FR2 ... (R^2 mod m)  #define FONE ... (R mod m)
 */
#include <stdint.h>
#include <string.h>

typedef unsigned __int128 u128;

static const uint64_t FNAME(MOD)[4] = FMOD;
static const uint64_t FNAME(R2)[4] = FR2;
static const uint64_t FNAME(ONE)[4] = FONE;

static inline int FNAME(is_zero)(const uint64_t a[4]) { return (a[0] | a[1] | a[2] | a[3]) == 0; }
static inline int FNAME(eq)(const uint64_t a[4], const uint64_t b[4]) {
    return ((a[0] ^ b[0]) | (a[1] ^ b[1]) | (a[2] ^ b[2]) | (a[3] ^ b[3])) == 0;
}
static inline void FNAME(set)(uint64_t r[4], const uint64_t a[4]) { memcpy(r, a, 32); }
static inline void FNAME(zero)(uint64_t r[4]) { r[0] = r[1] = r[2] = r[3] = 0; }
static inline void FNAME(one)(uint64_t r[4]) { memcpy(r, FNAME(ONE), 32); }

/* r = a - MOD if a >= MOD else a  (a < 2*MOD, carry = bit 256 of a) */
static inline void FNAME(cond_sub)(uint64_t r[4], const uint64_t a[4], uint64_t carry) {
    uint64_t t[4];
    u128 b = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - FNAME(MOD)[i] - (uint64_t)b;
        t[i] = (uint64_t)d;
        b = (d >> 64) & 1;
    }
    /* underflow (b=1) with no incoming carry means a < MOD: keep a */
    int keep = (b != 0) && (carry == 0);
    for (int i = 0; i < 4; i++) r[i] = keep ? a[i] : t[i];
}

static inline void FNAME(add)(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {
    uint64_t t[4];
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (u128)a[i] + b[i];
        t[i] = (uint64_t)c;
        c >>= 64;
    }
    FNAME(cond_sub)(r, t, (uint64_t)c);
}

static inline void FNAME(sub)(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {
    uint64_t t[4];
    u128 bw = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - b[i] - (uint64_t)bw;
        t[i] = (uint64_t)d;
        bw = (d >> 64) & 1;
    }
    if (bw) {
        u128 c = 0;
        for (int i = 0; i < 4; i++) {
            c += (u128)t[i] + FNAME(MOD)[i];
            t[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    memcpy(r, t, 32);
}

static inline void FNAME(neg)(uint64_t r[4], const uint64_t a[4]) {
    uint64_t z[4] = {0, 0, 0, 0};
    FNAME(sub)(r, z, a);
}

static inline void FNAME(dbl)(uint64_t r[4], const uint64_t a[4]) { FNAME(add)(r, a, a); }

/* Montgomery product, coarsely-integrated operand scanning: r = a*b*R^-1 mod MOD */
static inline void FNAME(mul)(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {
    uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
    const uint64_t* m = FNAME(MOD);
    for (int i = 0; i < 4; i++) {
        u128 acc;
        uint64_t c, bi = b[i], t5;
        acc = (u128)a[0] * bi + t0; t0 = (uint64_t)acc; c = (uint64_t)(acc >> 64);
        acc = (u128)a[1] * bi + t1 + c; t1 = (uint64_t)acc; c = (uint64_t)(acc >> 64);
        acc = (u128)a[2] * bi + t2 + c; t2 = (uint64_t)acc; c = (uint64_t)(acc >> 64);
        acc = (u128)a[3] * bi + t3 + c; t3 = (uint64_t)acc; c = (uint64_t)(acc >> 64);
        acc = (u128)t4 + c; t4 = (uint64_t)acc; t5 = (uint64_t)(acc >> 64);
        uint64_t q = t0 * (uint64_t)FINV;
        acc = (u128)q * m[0] + t0; c = (uint64_t)(acc >> 64);
        acc = (u128)q * m[1] + t1 + c; t0 = (uint64_t)acc; c = (uint64_t)(acc >> 64);
        acc = (u128)q * m[2] + t2 + c; t1 = (uint64_t)acc; c = (uint64_t)(acc >> 64);
        acc = (u128)q * m[3] + t3 + c; t2 = (uint64_t)acc; c = (uint64_t)(acc >> 64);
        acc = (u128)t4 + c; t3 = (uint64_t)acc; t4 = t5 + (uint64_t)(acc >> 64);
    }
    uint64_t t[4] = {t0, t1, t2, t3};
    FNAME(cond_sub)(r, t, t4);
}

static inline void FNAME(sqr)(uint64_t r[4], const uint64_t a[4]) { FNAME(mul)(r, a, a); }

static inline void FNAME(to_mont)(uint64_t r[4], const uint64_t a[4]) { FNAME(mul)(r, a, FNAME(R2)); }
static inline void FNAME(from_mont)(uint64_t r[4], const uint64_t a[4]) {
    uint64_t one[4] = {1, 0, 0, 0};
    FNAME(mul)(r, a, one);
}

/* r = a^e, e given as 4 little-endian u64 (plain integer) */
static inline void FNAME(pow)(uint64_t r[4], const uint64_t a[4], const uint64_t e[4]) {
    uint64_t acc[4], base[4];
    FNAME(one)(acc);
    FNAME(set)(base, a);
    for (int i = 0; i < 256; i++) {
        if ((e[i >> 6] >> (i & 63)) & 1) FNAME(mul)(acc, acc, base);
        FNAME(sqr)(base, base);
    }
    FNAME(set)(r, acc);
}

static inline void FNAME(pow_u64)(uint64_t r[4], const uint64_t a[4], uint64_t e) {
    uint64_t ee[4] = {e, 0, 0, 0};
    FNAME(pow)(r, a, ee);
}

/* Fermat inverse: a^(MOD-2); inverse of 0 is 0 (halo2's invert() returns CtOption::none there;
 * the batch-invert callers upstream skip zeros the same way). */
static inline void FNAME(inv)(uint64_t r[4], const uint64_t a[4]) {
    uint64_t e[4];
    memcpy(e, FNAME(MOD), 32);
    e[0] -= 2; /* both moduli end in ...01 / ...47: no borrow */
    FNAME(pow)(r, a, e);
}

/* Montgomery's trick over n elements; zeros stay zero (BatchInvert contract in `ff`). */
static inline void FNAME(batch_inv)(uint64_t* a, size_t n, uint64_t* scratch /* n*4 */) {
    uint64_t acc[4];
    FNAME(one)(acc);
    for (size_t i = 0; i < n; i++) {
        FNAME(set)(scratch + 4 * i, acc);
        if (!FNAME(is_zero)(a + 4 * i)) FNAME(mul)(acc, acc, a + 4 * i);
    }
    FNAME(inv)(acc, acc);
    for (size_t i = n; i-- > 0;) {
        if (FNAME(is_zero)(a + 4 * i)) continue;
        uint64_t t[4];
        FNAME(mul)(t, acc, scratch + 4 * i);
        FNAME(mul)(acc, acc, a + 4 * i);
        FNAME(set)(a + 4 * i, t);
    }
}
