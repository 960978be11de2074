/*
 * quotient.c — CPU oracle restatement of halo2_proofs::plonk::evaluation::Evaluator::evaluate_h
 * followed by EvaluationDomain::divide_by_vanishing_poly, specialised to the constraint system
 * that ECDSACircuit::configure builds (halo2-circuits/src/ecc/ecdsa_p256.rs:94-115: halo2-lib
 * FlexGate + Range).  TEST INFRASTRUCTURE ONLY (see zkw_oracle.h).
 *
 * The constraint list and its folding order are the ones the reference's generated verifier
 * re-evaluates at the challenge point (proving-server/P256Verifier.yul:406-547, k = 17):
 *   gates        yul:406-428   q_c * (a_c + a_c(wX) * a_c(w^2 X) - a_c(w^3 X)) per gate column
 *   permutation  yul:429-518   l_0 (1 - z_0); l_last (z_last^2 - z_last);
 *                              l_0 (z_i - z_{i-1}(w^last X)) for i >= 1;
 *                              l_active (z_i(wX) prod(v + beta sigma + gamma)
 *                                        - z_i(X) prod(v + delta^j beta X + gamma)) per set
 *   lookups      yul:519-547   l_0 (1 - z); l_last (z^2 - z);
 *                              l_active (z(wX)(a'+beta)(s'+gamma) - z (a+beta)(s+gamma));
 *                              l_0 (a' - s'); l_active (a' - s')(a' - a'(w^-1 X))
 * each folded as value = value * y + constraint, gates first.  With a single input / table
 * expression per lookup, theta-compression is the identity (theta is unused, as in yul:519-530).
 * Rotation r on the 2^k domain is an index shift of r * 2^(ext_k - k) on the extended coset.
 */
#include <omp.h>
#include <stdlib.h>
#include "bn254_internal.h"
#include "zkw_oracle.h"

int zko_quotient_ecdsa(const zkw_quotient_inputs* in, uint64_t* h_ext, int threads) {
    if (threads <= 0) threads = omp_get_max_threads();
    const zkw_circuit_shape* sh = &in->shape;
    const unsigned A = sh->num_advice, L = sh->num_lookup_advice, F = sh->num_fixed;
    const unsigned ncols = F + A + L;
    const unsigned chunk_len = sh->cs_degree - 2;
    const unsigned nsets = zkw_shape_perm_sets(sh);
    const unsigned nlk = zkw_shape_lookups(sh);
    if (sh->ext_k < sh->k || sh->ext_k - sh->k > 4 || chunk_len == 0) return ZKW_ERR_INVALID;
    if (L == 0 && (A != 1 || in->q_lookup == NULL)) return ZKW_ERR_UNSUPPORTED;
    const size_t en = (size_t)1 << sh->ext_k;
    const size_t mask = en - 1;
    const long rot_scale = 1L << (sh->ext_k - sh->k);
    const long last_rot = -((long)sh->blinding_factors + 1);

    zko_domain dom;
    zko_domain_new(&dom, sh->cs_degree, sh->k);
    if (dom.ext_k != sh->ext_k) return ZKW_ERR_INVALID;

    /* permutation column values in permutation order */
    const uint64_t** pcol = (const uint64_t**)malloc(sizeof(void*) * ncols);
    for (unsigned i = 0; i < F; i++) pcol[i] = in->constants[i];
    for (unsigned i = 0; i < A + L; i++) pcol[F + i] = in->advice[i];

    uint64_t one[4];
    fr_one(one);
    uint64_t delta_start[4]; /* beta * zeta */
    fr_mul(delta_start, in->beta, FR_ZETA_M);

#pragma omp parallel num_threads(threads)
    {
        int tid = omp_get_thread_num(), nt = omp_get_num_threads();
        size_t lo = en * (size_t)tid / (size_t)nt, hi = en * (size_t)(tid + 1) / (size_t)nt;
        uint64_t beta_term[4]; /* ext_omega^idx */
        fr_pow_u64(beta_term, dom.ext_omega, (uint64_t)lo);
        for (size_t idx = lo; idx < hi; idx++) {
            const size_t r1 = (idx + (size_t)rot_scale) & mask;
            const size_t r2 = (idx + (size_t)(2 * rot_scale)) & mask;
            const size_t r3 = (idx + (size_t)(3 * rot_scale)) & mask;
            const size_t rprev = (idx + en - (size_t)rot_scale) & mask;
            const size_t rlast = (size_t)((long)idx + last_rot * rot_scale + (long)en * 16) & mask;
            uint64_t v[4], t[4], u[4];
            fr_zero(v);
#define FOLD(c) do { fr_mul(v, v, in->y); fr_add(v, v, (c)); } while (0)
            /* gates */
            for (unsigned c = 0; c < A; c++) {
                const uint64_t* a = in->advice[c];
                fr_mul(t, a + 4 * r1, a + 4 * r2);
                fr_add(t, t, a + 4 * idx);
                fr_sub(t, t, a + 4 * r3);
                fr_mul(t, t, in->q_enable[c] + 4 * idx);
                FOLD(t);
            }
            const uint64_t* l0 = in->l0 + 4 * idx;
            const uint64_t* ll = in->l_last + 4 * idx;
            const uint64_t* la = in->l_active + 4 * idx;
            /* permutation */
            if (nsets) {
                const uint64_t* z0 = in->perm_z[0] + 4 * idx;
                fr_sub(t, one, z0); fr_mul(t, t, l0); FOLD(t);
                const uint64_t* zl = in->perm_z[nsets - 1] + 4 * idx;
                fr_sqr(t, zl); fr_sub(t, t, zl); fr_mul(t, t, ll); FOLD(t);
                for (unsigned s = 1; s < nsets; s++) {
                    fr_sub(t, in->perm_z[s] + 4 * idx, in->perm_z[s - 1] + 4 * rlast);
                    fr_mul(t, t, l0);
                    FOLD(t);
                }
                uint64_t cur_delta[4];
                fr_mul(cur_delta, delta_start, beta_term);
                for (unsigned s = 0; s < nsets; s++) {
                    unsigned c0 = s * chunk_len, c1 = c0 + chunk_len > ncols ? ncols : c0 + chunk_len;
                    uint64_t left[4], right[4];
                    fr_set(left, in->perm_z[s] + 4 * r1);
                    for (unsigned c = c0; c < c1; c++) {
                        fr_mul(t, in->beta, in->sigma[c] + 4 * idx);
                        fr_add(t, t, pcol[c] + 4 * idx);
                        fr_add(t, t, in->gamma);
                        fr_mul(left, left, t);
                    }
                    fr_set(right, in->perm_z[s] + 4 * idx);
                    for (unsigned c = c0; c < c1; c++) {
                        fr_add(t, pcol[c] + 4 * idx, cur_delta);
                        fr_add(t, t, in->gamma);
                        fr_mul(right, right, t);
                        fr_mul(cur_delta, cur_delta, FR_DELTA_M);
                    }
                    fr_sub(t, left, right);
                    fr_mul(t, t, la);
                    FOLD(t);
                }
            }
            /* lookups */
            for (unsigned k = 0; k < nlk; k++) {
                const uint64_t* z = in->lookup_z[k];
                const uint64_t* ap = in->lookup_a[k];
                const uint64_t* sp = in->lookup_s[k];
                uint64_t inp[4];
                if (L) fr_set(inp, in->advice[A + k] + 4 * idx);
                else fr_mul(inp, in->q_lookup + 4 * idx, in->advice[0] + 4 * idx);
                fr_sub(t, one, z + 4 * idx); fr_mul(t, t, l0); FOLD(t);
                fr_sqr(t, z + 4 * idx); fr_sub(t, t, z + 4 * idx); fr_mul(t, t, ll); FOLD(t);
                /* z(wX)(a'+beta)(s'+gamma) - z(X)(a+beta)(s+gamma) */
                fr_add(t, ap + 4 * idx, in->beta);
                fr_add(u, sp + 4 * idx, in->gamma);
                fr_mul(t, t, u);
                fr_mul(t, t, z + 4 * r1);
                uint64_t w[4];
                fr_add(w, inp, in->beta);
                fr_add(u, in->table + 4 * idx, in->gamma);
                fr_mul(w, w, u);
                fr_mul(w, w, z + 4 * idx);
                fr_sub(t, t, w);
                fr_mul(t, t, la);
                FOLD(t);
                uint64_t ams[4];
                fr_sub(ams, ap + 4 * idx, sp + 4 * idx);
                fr_mul(t, ams, l0); FOLD(t);
                fr_sub(t, ap + 4 * idx, ap + 4 * rprev);
                fr_mul(t, t, ams);
                fr_mul(t, t, la);
                FOLD(t);
            }
#undef FOLD
            /* divide_by_vanishing_poly */
            fr_mul(h_ext + 4 * idx, v, dom.t_evaluations[idx & (size_t)(rot_scale - 1)]);
            fr_mul(beta_term, beta_term, dom.ext_omega);
        }
    }
    free(pcol);
    return ZKW_OK;
}
