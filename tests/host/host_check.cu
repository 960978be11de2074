// host_check.cu — runs the product's __host__ __device__ field / curve formulas on the CPU and
// compares them with the CPU oracle (test infrastructure; built and run by tests/test_host_formulas.py).
#include <cstdio>
#include <cstring>
#include <vector>
#include "curve.cuh"
#include "inverse.cuh"
#include "../../oracle/zkw_oracle.h"
using namespace zkw;

static int fails = 0;
#define CHECK(c, ...) do { if (!(c)) { fails++; printf("FAIL %s:%d ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } } while (0)

template <class F> static F from_u64(const uint64_t* p) { F r; memcpy(r.l, p, 32); return r; }

int main() {
    const int N = 2000;
    std::vector<uint64_t> a(4 * N), b(4 * N);
    zko_fr_random(a.data(), N, 11);
    zko_fr_random(b.data(), N, 12);
    // Fr and Fq field ops (random Fr values are < r < p so they are valid Fq elements too)
    for (int i = 0; i < N; i++) {
        uint64_t r[4];
        Fr x = from_u64<Fr>(&a[4 * i]), y = from_u64<Fr>(&b[4 * i]);
        Fr z = x * y; zko_fr_mul(r, &a[4 * i], &b[4 * i]); CHECK(!memcmp(z.l, r, 32), "fr mul %d", i);
        z = x + y; zko_fr_add(r, &a[4 * i], &b[4 * i]); CHECK(!memcmp(z.l, r, 32), "fr add %d", i);
        z = x - y; zko_fr_sub(r, &a[4 * i], &b[4 * i]); CHECK(!memcmp(z.l, r, 32), "fr sub %d", i);
        Fq u = from_u64<Fq>(&a[4 * i]), v = from_u64<Fq>(&b[4 * i]);
        Fq w = u * v; zko_fq_mul(r, &a[4 * i], &b[4 * i]); CHECK(!memcmp(w.l, r, 32), "fq mul %d", i);
        w = u + v; zko_fq_add(r, &a[4 * i], &b[4 * i]); CHECK(!memcmp(w.l, r, 32), "fq add %d", i);
        w = u - v; zko_fq_sub(r, &a[4 * i], &b[4 * i]); CHECK(!memcmp(w.l, r, 32), "fq sub %d", i);
        if (i < 20) {
            z = x.inv(); zko_fr_inv(r, &a[4 * i]); CHECK(!memcmp(z.l, r, 32), "fr inv %d", i);
            w = u.inv(); zko_fq_inv(r, &a[4 * i]); CHECK(!memcmp(w.l, r, 32), "fq inv %d", i);
            z = x.from_mont(); zko_fr_from_mont(r, &a[4 * i]); CHECK(!memcmp(z.l, r, 32), "fr from_mont %d", i);
            z = x.to_mont(); zko_fr_to_mont(r, &a[4 * i]); CHECK(!memcmp(z.l, r, 32), "fr to_mont %d", i);
        }
    }
    // curve: random points = scalars * G
    const int M = 64;
    std::vector<uint64_t> pts(8 * M);
    zko_g1_fixed_base_mul(pts.data(), a.data(), M, 1);
    auto affine_of = [](const G1Xyzz& p, uint64_t out[8]) {
        Fq jx, jy, jz; p.to_jacobian(jx, jy, jz);
        uint64_t xyz[12]; memcpy(xyz, jx.l, 32); memcpy(xyz + 4, jy.l, 32); memcpy(xyz + 8, jz.l, 32);
        zko_g1_to_affine(out, xyz);
    };
    G1Xyzz acc = G1Xyzz::identity();
    uint64_t oacc[12] = {0}; oacc[4] = 1; // identity (z = 0)
    for (int i = 0; i < M; i++) {
        G1Affine p = G1Affine::load(&pts[8 * i]);
        bool neg = (i % 5) == 3;
        acc.add_mixed(p, neg);
        uint64_t q[8]; memcpy(q, &pts[8 * i], 64);
        if (neg) { uint64_t z[4] = {0, 0, 0, 0}; zko_fq_sub(q + 4, z, q + 4); }
        zko_g1_add_mixed(oacc, oacc, q);
        uint64_t g[8], o[8]; affine_of(acc, g); zko_g1_to_affine(o, oacc);
        CHECK(!memcmp(g, o, 64), "add_mixed step %d", i);
    }
    // doubling via add_mixed of the same point, cancellation, general add, dbl
    {
        G1Affine p = G1Affine::load(&pts[0]);
        G1Xyzz t = G1Xyzz::from_affine(p); t.add_mixed(p);
        uint64_t o12[12], pj[12] = {0}; memcpy(pj, &pts[0], 64); zko_fq_to_mont(pj + 8, (const uint64_t[4]){1, 0, 0, 0});
        zko_g1_double(o12, pj);
        uint64_t g[8], o[8]; affine_of(t, g); zko_g1_to_affine(o, o12);
        CHECK(!memcmp(g, o, 64), "mixed doubling");
        G1Xyzz t2 = t.dbl(); zko_g1_double(o12, o12); affine_of(t2, g); zko_g1_to_affine(o, o12);
        CHECK(!memcmp(g, o, 64), "xyzz dbl");
        G1Xyzz c = G1Xyzz::from_affine(p); c.add_mixed(p, true);
        CHECK(c.is_identity(), "P + (-P)");
        G1Xyzz s = acc; s.add(t2);
        uint64_t sum12[12]; zko_g1_add(sum12, oacc, o12); affine_of(s, g); zko_g1_to_affine(o, sum12);
        CHECK(!memcmp(g, o, 64), "xyzz add");
        G1Xyzz d = t2; d.add(t2); zko_g1_double(o12, o12); affine_of(d, g); zko_g1_to_affine(o, o12);
        CHECK(!memcmp(g, o, 64), "xyzz add equal -> dbl");
        G1Xyzz e = G1Xyzz::identity(); e.add(t2); e.add(G1Xyzz::identity());
        affine_of(e, g); affine_of(t2, o); CHECK(!memcmp(g, o, 64), "identity handling");
    }
    // binary-GCD inversion (inverse.cuh) against the oracle's inversion: random elements and the edge cases of the
    // divstep loop (0, 1, -1, 2, single bits in every limb, values just below the modulus)
    {
        const int M2 = 3000;
        std::vector<uint64_t> v(4 * M2);
        zko_fr_random(v.data(), M2, 99);
        for (int i = 0; i < M2; i++) {
            uint64_t r[4];
            Fr x = from_u64<Fr>(&v[4 * i]);
            Fq y = from_u64<Fq>(&v[4 * i]);
            if (i == 0) { x = Fr::zero(); y = Fq::zero(); }
            if (i == 1) { x = Fr::one(); y = Fq::one(); }
            if (i == 2) { x = Fr::zero() - Fr::one(); y = Fq::zero() - Fq::one(); }
            if (i == 3) { x = Fr::one() + Fr::one(); y = Fq::one() + Fq::one(); }
            if (i >= 4 && i < 36) { x = Fr::zero(); x.l[(i - 4) / 4] = 1u << ((i * 7) % 32); y = Fq::zero(); y.l[(i - 4) / 4] = 1u << ((i * 11) % 32); }
            if (i >= 36 && i < 44) { x = Fr::zero() - from_u64<Fr>(&v[4 * i]).from_mont().from_mont(); x = Fr::zero() - Fr::one() - Fr::one(); x.l[0] -= (i - 36); y = Fq::zero() - Fq::one(); y.l[0] -= (i - 36); }
            Fr xi = fp_inv_bingcd(x);
            zko_fr_inv(r, (const uint64_t*)x.l); CHECK(!memcmp(xi.l, r, 32), "fr bingcd inv %d", i);
            Fq yi = fp_inv_bingcd(y);
            zko_fq_inv(r, (const uint64_t*)y.l); CHECK(!memcmp(yi.l, r, 32), "fq bingcd inv %d", i);
        }
    }
    printf(fails ? "HOST_CHECK FAILED (%d)\n" : "HOST_CHECK OK\n", fails);
    return fails ? 1 : 0;
}
