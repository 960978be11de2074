/*
 * msm.c — CPU oracle restatement of halo2_proofs::arithmetic::{best_multiexp, multiexp_serial}.
 * TEST INFRASTRUCTURE ONLY (see zkw_oracle.h).
 *
 * The reference reaches this through ParamsKZG::commit / commit_lagrange inside create_proof
 * (halo2-circuits/src/ecc/ecdsa_p256.rs:366-373, 416-423, 555-562) and keygen (:259-260).  The
 * crate is an un-vendored dependency (Cargo.toml:12-13), so the algorithm below is restated from
 * its published definition:
 *   - scalars are taken out of Montgomery form first (upstream: `to_repr()`);
 *   - window c = 1 if n < 4, 3 if n < 32, else ceil(ln n); segments = 256/c + 1;
 *   - per segment, from the top: c doublings of the accumulator, 2^c - 1 buckets filled by
 *     mixed additions, running-sum ("summation by parts") reduction from the top bucket down;
 *   - best_multiexp splits the input into one contiguous chunk of n / num_threads per thread,
 *     runs the serial routine on each, and folds the partial results in chunk order.
 * The result is a projective point; only its affine normalisation is canonical.
 */
#include <math.h>
#include <omp.h>
#include <stdlib.h>
#include "bn254_internal.h"
#include "zkw_oracle.h"

static inline unsigned get_at(unsigned segment, unsigned c, const uint64_t repr[4]) {
    unsigned skip_bits = segment * c;
    if (skip_bits >= 256) return 0;
    unsigned limb = skip_bits >> 6, off = skip_bits & 63;
    uint64_t v = repr[limb] >> off;
    if (off + c > 64 && limb + 1 < 4) v |= repr[limb + 1] << (64 - off);
    return (unsigned)(v & ((1ULL << c) - 1));
}

static void multiexp_serial(const uint64_t* reprs /* canonical */, const g1a_t* bases, size_t n, g1_t* acc) {
    unsigned c;
    if (n < 4) c = 1;
    else if (n < 32) c = 3;
    else c = (unsigned)ceil(log((double)n));
    unsigned segments = 256 / c + 1;
    size_t nb = ((size_t)1 << c) - 1;
    g1_t* buckets = (g1_t*)malloc(sizeof(g1_t) * nb);
    for (unsigned seg = segments; seg-- > 0;) {
        for (unsigned i = 0; i < c; i++) g1_double(acc, acc);
        for (size_t b = 0; b < nb; b++) g1_set_identity(&buckets[b]);
        for (size_t i = 0; i < n; i++) {
            unsigned d = get_at(seg, c, reprs + 4 * i);
            if (d) g1_add_mixed(&buckets[d - 1], &buckets[d - 1], &bases[i]);
        }
        g1_t running;
        g1_set_identity(&running);
        for (size_t b = nb; b-- > 0;) {
            g1_add(&running, &running, &buckets[b]);
            g1_add(acc, acc, &running);
        }
    }
    free(buckets);
}

void zko_best_multiexp(uint64_t out_xyz[12], const uint64_t* scalars, const uint64_t* bases_xy, size_t n, int threads) {
    if (threads <= 0) threads = omp_get_max_threads();
    g1_t total;
    g1_set_identity(&total);
    if (n == 0) { memcpy(out_xyz, &total, 96); return; }
    uint64_t* reprs = (uint64_t*)malloc(n * 32);
    zko_fr_vec_from_mont(reprs, scalars, n);
    const g1a_t* bases = (const g1a_t*)bases_xy;
    if (n > (size_t)threads) {
        size_t chunk = n / (size_t)threads;
        size_t num_chunks = (n + chunk - 1) / chunk;
        g1_t* results = (g1_t*)malloc(sizeof(g1_t) * num_chunks);
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
        for (size_t ci = 0; ci < num_chunks; ci++) {
            size_t lo = ci * chunk, hi = lo + chunk > n ? n : lo + chunk;
            g1_set_identity(&results[ci]);
            multiexp_serial(reprs + 4 * lo, bases + lo, hi - lo, &results[ci]);
        }
        for (size_t ci = 0; ci < num_chunks; ci++) g1_add(&total, &total, &results[ci]);
        free(results);
    } else {
        multiexp_serial(reprs, bases, n, &total);
    }
    free(reprs);
    memcpy(out_xyz, &total, 96);
}

void zko_msm_naive(uint64_t out_xyz[12], const uint64_t* scalars, const uint64_t* bases_xy, size_t n) {
    g1_t total;
    g1_set_identity(&total);
    for (size_t i = 0; i < n; i++) {
        g1_t t;
        zko_g1_mul((uint64_t*)&t, bases_xy + 8 * i, scalars + 4 * i);
        g1_add(&total, &total, &t);
    }
    memcpy(out_xyz, &total, 96);
}
