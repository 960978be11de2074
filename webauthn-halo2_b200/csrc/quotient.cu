// quotient.cu — coset-domain quotient evaluation for the P-256 ECDSA circuit on sm_100a: the device
// replacement of halo2_proofs::plonk::evaluation::Evaluator::evaluate_h followed by
// EvaluationDomain::divide_by_vanishing_poly, reached from the reference through create_proof
// (halo2-circuits/src/ecc/ecdsa_p256.rs:366-373, 416-423, 555-562).
//
// The constraint system is the one ECDSACircuit::configure builds (ecdsa_p256.rs:94-115, halo2-lib
// FlexGate + Range) and the reference's generated verifier re-evaluates at the challenge point
// (proving-server/P256Verifier.yul:406-547):
//   gates        q_c (a_c + a_c(wX) a_c(w^2 X) - a_c(w^3 X))                     per gate column
//   permutation  l_0 (1 - z_0);  l_last (z_last^2 - z_last);  l_0 (z_i - z_{i-1}(w^last X)), i >= 1;
//                l_active (z_i(wX) prod(v + beta sigma + gamma) - z_i prod(v + delta^j beta X + gamma))
//   lookups      l_0 (1 - z);  l_last (z^2 - z);
//                l_active (z(wX)(a'+beta)(s'+gamma) - z (a+beta)(s+gamma));
//                l_0 (a' - s');  l_active (a' - s')(a' - a'(w^-1 X))
// folded as h = h*y + constraint in that order, then multiplied by 1/(X^n - 1) on the coset.  The fold is evaluated
// grouped by the Lagrange factor: every constraint after the gates is l_0, l_last or l_active times something, so
//   h = gates(y) * y^m + l_0 * sum_j y^e_j g_j + l_last * sum_j y^e_j g_j + l_active * sum_j y^e_j g_j
// with the powers of y precomputed on the host - one product per constraint plus three, instead of two per constraint
// (the same field element: exact arithmetic; 32 products per row instead of 39 at k = 19).
//
// One thread per extended-domain row: every input coset is read once per row (rotated reads of the
// same column land on neighbouring rows' sectors and are served by L1/L2), the running value never
// leaves registers, and h is written once.  Algorithmic traffic: 32 B x (distinct input cosets + 1)
// per row.  The per-row arithmetic (~40 Montgomery products at k = 19) makes this ALU bound.
#include "common.cuh"

namespace zkw {

struct QuotArgs {
    int A, L, F, ncols, chunk_len, nsets, nlk;
    int ext_k;
    unsigned rot_scale;      // 2^(ext_k - k)
    int last_rot;            // -(blinding_factors + 1)
    const uint4* const* advice;
    const uint4* const* constants;
    const uint4* const* q_enable;
    const uint4* const* sigma;
    const uint4* const* perm_z;
    const uint4* const* lookup_z;
    const uint4* const* lookup_a;
    const uint4* const* lookup_s;
    const uint4* table;
    const uint4* q_lookup;
    const uint4* l0;
    const uint4* l_last;
    const uint4* l_active;
    const uint4* ext_tw;     // ext_omega^i, i < 2^(ext_k-1)
    const uint4* ypow;       // y^0 .. y^rest_constraints
    int rest_constraints;    // constraints after the gates: 2 + (nsets - 1) + nsets + 5 * nlk
    uint4* h;
    Fr y, beta, gamma, beta_zeta, delta, one;
    Fr t_evals[16];
};

__global__ void __launch_bounds__(128) quotient_kernel(const QuotArgs q) {
    const size_t en = (size_t)1 << q.ext_k;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= en) return;
    const size_t mask = en - 1;
    const size_t r1 = (idx + q.rot_scale) & mask;
    const size_t r2 = (idx + 2 * (size_t)q.rot_scale) & mask;
    const size_t r3 = (idx + 3 * (size_t)q.rot_scale) & mask;
    const size_t rprev = (idx + en - q.rot_scale) & mask;
    const size_t rlast = (idx + en * 16 + (size_t)((long long)q.last_rot * (long long)q.rot_scale)) & mask;

    const Fr y = q.y, beta = q.beta, gamma = q.gamma;
    Fr v = Fr::zero();
    // gates
    for (int c = 0; c < q.A; c++) {
        const uint4* a = q.advice[c];
        Fr t = Fr::load_nc(a + 2 * r1) * Fr::load_nc(a + 2 * r2);
        t = t + Fr::load_nc(a + 2 * idx) - Fr::load_nc(a + 2 * r3);
        t = t * Fr::load_nc(q.q_enable[c] + 2 * idx);
        v = c ? v * y + t : t;
    }
    const Fr l0 = Fr::load_nc(q.l0 + 2 * idx);
    const Fr ll = Fr::load_nc(q.l_last + 2 * idx);
    const Fr la = Fr::load_nc(q.l_active + 2 * idx);
    // the constraints after the gates, each weighted by its power of y and summed per Lagrange factor
    Fr acc0 = Fr::zero(), accl = Fr::zero(), acca = Fr::zero();
    int e = q.rest_constraints - 1;
    auto weighted = [&](const Fr& g) -> Fr {
        const Fr r = e > 0 ? g * Fr::load_nc(q.ypow + 2 * e) : g;
        e--;
        return r;
    };
    // permutation
    if (q.nsets) {
        {
            Fr z0 = Fr::load_nc(q.perm_z[0] + 2 * idx);
            acc0 = acc0 + weighted(q.one - z0);
            Fr zl = Fr::load_nc(q.perm_z[q.nsets - 1] + 2 * idx);
            accl = accl + weighted(zl.sqr() - zl);
        }
        for (int s = 1; s < q.nsets; s++) {
            Fr t = Fr::load_nc(q.perm_z[s] + 2 * idx) - Fr::load_nc(q.perm_z[s - 1] + 2 * rlast);
            acc0 = acc0 + weighted(t);
        }
        // beta * zeta * ext_omega^idx
        const size_t half = en >> 1;
        Fr x = Fr::load_nc(q.ext_tw + 2 * (idx & (half - 1)));
        if (idx >= half) x = x.neg();
        Fr cur_delta = q.beta_zeta * x;
        for (int s = 0; s < q.nsets; s++) {
            const int c0 = s * q.chunk_len;
            const int c1 = min(c0 + q.chunk_len, q.ncols);
            Fr left = Fr::load_nc(q.perm_z[s] + 2 * r1);
            Fr right = Fr::load_nc(q.perm_z[s] + 2 * idx);
            for (int c = c0; c < c1; c++) {
                const uint4* col = c < q.F ? q.constants[c] : q.advice[c - q.F];
                Fr val = Fr::load_nc(col + 2 * idx);
                Fr sg = Fr::load_nc(q.sigma[c] + 2 * idx);
                left = left * (beta * sg + val + gamma);
                right = right * (val + cur_delta + gamma);
                if (c + 1 < q.ncols) cur_delta = cur_delta * q.delta;
            }
            acca = acca + weighted(left - right);
        }
    }
    // lookups
    for (int k = 0; k < q.nlk; k++) {
        const uint4* zc = q.lookup_z[k];
        const Fr z = Fr::load_nc(zc + 2 * idx);
        const Fr zn = Fr::load_nc(zc + 2 * r1);
        const Fr ap = Fr::load_nc(q.lookup_a[k] + 2 * idx);
        const Fr app = Fr::load_nc(q.lookup_a[k] + 2 * rprev);
        const Fr sp = Fr::load_nc(q.lookup_s[k] + 2 * idx);
        Fr inp;
        if (q.L) inp = Fr::load_nc(q.advice[q.A + k] + 2 * idx);
        else inp = Fr::load_nc(q.q_lookup + 2 * idx) * Fr::load_nc(q.advice[0] + 2 * idx);
        const Fr tab = Fr::load_nc(q.table + 2 * idx);
        // the five lookup constraints carry the weights y^e .. y^(e-4); the two l_0 terms and the two l_active terms are each
        // folded as ONE two-product Montgomery pass (Fr::dot_lazy: one reduction per pair)
        const Fr lhs = (ap + beta) * (sp + gamma) * zn;
        const Fr rhs = (inp + beta) * (tab + gamma) * z;
        const Fr ams = ap - sp;
        {
            const Fr g0[2] = {q.one - z, ams}, w0[2] = {Fr::load_nc(q.ypow + 2 * e), Fr::load_nc(q.ypow + 2 * (e - 3))};
            acc0 = acc0 + Fr::dot_lazy<2>(g0, w0).normalized();
            accl = accl + (z.sqr() - z) * Fr::load_nc(q.ypow + 2 * (e - 1));
            const Fr ga[2] = {lhs - rhs, (ap - app) * ams}, wa[2] = {Fr::load_nc(q.ypow + 2 * (e - 2)), Fr::load_nc(q.ypow + 2 * (e - 4))};
            acca = acca + Fr::dot_lazy<2>(ga, wa).normalized();
            e -= 5;
        }
    }
    {   // h = gates * y^rest + acc0 l_0 + accl l_last + acca l_active: four products, one reduction
        const Fr fa[4] = {v, acc0, accl, acca}, fb[4] = {Fr::load_nc(q.ypow + 2 * q.rest_constraints), l0, ll, la};
        v = Fr::dot_lazy<4>(fa, fb).normalized();
    }
    v = v * q.t_evals[idx & (q.rot_scale - 1)];
    v.store(q.h + 2 * idx);
}

static Fr fr_host(const uint64_t v[4]) {
    Fr r;
    memcpy(r.l, v, 32);
    return r;
}

int quotient_run(zkw_ctx* ctx, const zkw_quotient_inputs* in, uint64_t* h_ext_dev) {
    const zkw_circuit_shape& sh = in->shape;
    const unsigned A = sh.num_advice, L = sh.num_lookup_advice, F = sh.num_fixed;
    if (sh.cs_degree < 3 || sh.ext_k < sh.k || sh.ext_k - sh.k > 4 || sh.ext_k > 28 || A == 0 || F == 0) return ZKW_ERR_INVALID;
    if (L == 0 && (A != 1 || !in->q_lookup)) return ZKW_ERR_UNSUPPORTED;
    {
        unsigned ek = sh.k;
        while ((1ull << ek) < (1ull << sh.k) * (sh.cs_degree - 1)) ek++;
        if (ek != sh.ext_k) return ZKW_ERR_INVALID;
    }
    const unsigned ncols = F + A + L;
    const unsigned nsets = zkw_shape_perm_sets(&sh);
    const unsigned nlk = zkw_shape_lookups(&sh);
    if (!in->advice || !in->constants || !in->q_enable || !in->sigma || !in->perm_z || !in->lookup_z || !in->lookup_a ||
        !in->lookup_s || !in->table || !in->l0 || !in->l_last || !in->l_active || !h_ext_dev)
        return ZKW_ERR_INVALID;

    // flatten the pointer tables and ship them to the device
    std::vector<const uint64_t*> flat;
    auto push = [&](const uint64_t* const* t, unsigned cnt) { size_t o = flat.size(); for (unsigned i = 0; i < cnt; i++) flat.push_back(t[i]); return o; };
    const size_t o_adv = push(in->advice, A + L);
    const size_t o_const = push(in->constants, F);
    const size_t o_q = push(in->q_enable, A);
    const size_t o_sig = push(in->sigma, ncols);
    const size_t o_pz = push(in->perm_z, nsets);
    const size_t o_lz = push(in->lookup_z, nlk);
    const size_t o_la = push(in->lookup_a, nlk);
    const size_t o_ls = push(in->lookup_s, nlk);
    for (auto p : flat) if (!p) return ZKW_ERR_INVALID;
    // powers of y for the grouped fold, stored behind the pointer table (32-byte aligned)
    const unsigned rest = 2 + (nsets - 1) + nsets + 5 * nlk;
    const size_t ypow_off = (flat.size() * sizeof(void*) + 31) & ~(size_t)31;
    std::vector<uint8_t> blob(ypow_off + ((size_t)rest + 1) * 32, 0);
    memcpy(blob.data(), flat.data(), flat.size() * sizeof(void*));
    {
        const Fr yv = fr_host(in->y);
        Fr pw = Fr::one();
        for (unsigned i = 0; i <= rest; i++) { memcpy(blob.data() + ypow_off + 32 * (size_t)i, pw.l, 32); pw = pw * yv; }
    }
    ZKW_TRY(ensure_buffer(ctx, ctx->ptr_table, blob.size()));
    ZKW_CUDA(ctx, cudaMemcpyAsync(ctx->ptr_table.ptr, blob.data(), blob.size(), cudaMemcpyHostToDevice, ctx->stream));
    // the copy reads pageable host memory: it is staged before the call returns, so `flat` may die
    const uint4* const* dev_tab = (const uint4* const*)ctx->ptr_table.ptr;

    DomainConsts dc;
    domain_consts(sh.k, sh.ext_k, &dc);
    const uint64_t* ext_tw = nullptr;
    ZKW_TRY(ntt_get_twiddles(ctx, dc.ext_omega, sh.ext_k, &ext_tw));

    QuotArgs q;
    memset(&q, 0, sizeof(q));
    q.A = (int)A; q.L = (int)L; q.F = (int)F; q.ncols = (int)ncols;
    q.chunk_len = (int)sh.cs_degree - 2; q.nsets = (int)nsets; q.nlk = (int)nlk;
    q.ext_k = (int)sh.ext_k;
    q.rot_scale = 1u << (sh.ext_k - sh.k);
    q.last_rot = -((int)sh.blinding_factors + 1);
    q.advice = dev_tab + o_adv; q.constants = dev_tab + o_const; q.q_enable = dev_tab + o_q; q.sigma = dev_tab + o_sig;
    q.perm_z = dev_tab + o_pz; q.lookup_z = dev_tab + o_lz; q.lookup_a = dev_tab + o_la; q.lookup_s = dev_tab + o_ls;
    q.table = (const uint4*)in->table; q.q_lookup = (const uint4*)in->q_lookup;
    q.l0 = (const uint4*)in->l0; q.l_last = (const uint4*)in->l_last; q.l_active = (const uint4*)in->l_active;
    q.ext_tw = (const uint4*)ext_tw;
    q.ypow = (const uint4*)((const uint8_t*)ctx->ptr_table.ptr + ypow_off);
    q.rest_constraints = (int)rest;
    q.h = (uint4*)h_ext_dev;
    q.y = fr_host(in->y); q.beta = fr_host(in->beta); q.gamma = fr_host(in->gamma);
    q.beta_zeta = q.beta * fr_host(dc.zeta);
    static const uint64_t delta_m[4] = {0x9a0c322befd78855ULL, 0x46e82d14249b563cULL, 0x5983a663e0b0b7a7ULL, 0x22ab452baaa111adULL};
    q.delta = fr_host(delta_m);
    q.one = Fr::one();
    for (unsigned i = 0; i < q.rot_scale; i++) q.t_evals[i] = fr_host(dc.t_evals[i]);
    const size_t en = (size_t)1 << sh.ext_k;
    { ProfScope ps_(ctx, "quotient_kernel"); quotient_kernel<<<(unsigned)((en + 127) / 128), 128, 0, ctx->stream>>>(q); }
    ZKW_LAUNCHED(ctx);
    return ZKW_OK;
}

}  // namespace zkw
