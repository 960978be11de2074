// srs.cu — development SRS generation on the device: the replacement of halo2-lib's `gen_srs(k)` /
// ParamsKZG::setup(k, rng) as called by the reference at halo2-circuits/src/ecc/ecdsa_p256.rs:258,
// 279, 338, 388, 430.  Upstream draws tau from an RNG and computes g[i] = tau^i * G followed by a
// group FFT for the Lagrange basis; here tau is an explicit argument (tests and the verifier need
// it reproducible) and both bases come from fixed-base scalar multiplications:
//     g[i]          = tau^i * G
//     g_lagrange[i] = L_i(tau) * G,   L_i(tau) = (tau^n - 1)/n * omega^i / (tau - omega^i)
// which is the same point set the group FFT yields (canonical affine form).
//
// Fixed-base multiplication: 32 byte-windows, table T[w][d] = d * 2^(8w) * G (8192 affine points,
// built once on the host with the shared field code), 32 mixed additions per scalar, one inversion
// per point to normalise.  Setup-time code: clarity over speed.
#include "common.cuh"

namespace zkw {

__global__ void __launch_bounds__(128) fixed_base_kernel(const uint4* __restrict__ scalars, const uint4* __restrict__ table,
                                                         uint4* __restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr s = Fr::load(scalars + 2 * i).from_mont();
    G1Xyzz acc = G1Xyzz::identity();
    for (int w = 0; w < 32; w++) {
        const uint32_t d = (s.l[w >> 2] >> ((w & 3) * 8)) & 0xffu;
        if (d) {
            G1Affine p = G1Affine::load_nc(table + 4 * (size_t)(w * 256 + d));
            acc.add_mixed(p);
        }
    }
    G1Affine a;
    if (acc.is_identity()) { a.x = Fq::zero(); a.y = Fq::zero(); }
    else {
        Fq tinv = fp_inv_bingcd(acc.zz * acc.zzz);
        a.x = acc.x * (acc.zzz * tinv);
        a.y = acc.y * (acc.zz * tinv);
    }
    a.store(out + 4 * i);
}

// out[i] = base^i for i < n (same scheme as the twiddle table)
__global__ void powers_kernel(uint4* out, Fr base, size_t n, unsigned run) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t start = t * run;
    if (start >= n) return;
    Fr w = base.pow((uint64_t)start);
    const size_t end = start + run < n ? start + run : n;
    for (size_t i = start; i < end; i++) {
        w.store(out + 2 * i);
        w = w * base;
    }
}

// out[i] = c * omega^i / (tau - omega^i), c = (tau^n - 1)/n ; omega^i read from pw[i]
__global__ void lagrange_scalars_kernel(const uint4* __restrict__ pw, uint4* __restrict__ out, Fr tau, Fr c, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr w = Fr::load(pw + 2 * i);
    Fr d = fp_inv_bingcd(tau - w);
    (c * w * d).store(out + 2 * i);
}

static std::vector<uint64_t> build_fixed_table_host() {
    // T[w][d] = d * 2^(8w) * G in affine form; XYZZ accumulation + one batched inversion
    const int W = 32, D = 256;
    std::vector<G1Xyzz> pts((size_t)W * D);
    G1Affine g;
    g.x = Fq::one();
    g.y = Fq::one() + Fq::one();
    G1Xyzz base = G1Xyzz::from_affine(g);
    for (int w = 0; w < W; w++) {
        G1Xyzz* row = &pts[(size_t)w * D];
        row[0] = G1Xyzz::identity();
        for (int d = 1; d < D; d++) { row[d] = row[d - 1]; row[d].add(base); }
        for (int i = 0; i < 8; i++) base = base.dbl();
    }
    std::vector<Fq> pref(pts.size());
    Fq run = Fq::one();
    for (size_t i = 0; i < pts.size(); i++) {
        pref[i] = run;
        if (!pts[i].is_identity()) run = run * (pts[i].zz * pts[i].zzz);
    }
    Fq inv = fp_inv_bingcd(run);
    std::vector<uint64_t> out(pts.size() * 8, 0);
    for (size_t i = pts.size(); i-- > 0;) {
        if (pts[i].is_identity()) continue;
        Fq tinv = inv * pref[i];
        inv = inv * (pts[i].zz * pts[i].zzz);
        Fq x = pts[i].x * (pts[i].zzz * tinv), y = pts[i].y * (pts[i].zz * tinv);
        memcpy(&out[i * 8], x.l, 32);
        memcpy(&out[i * 8 + 4], y.l, 32);
    }
    return out;
}

int fixed_base_table(zkw_ctx* ctx, const uint64_t** out_dev) {
    static DeviceBuffer tables[64];  // per device
    DeviceBuffer& t = tables[ctx->device & 63];
    if (!t.ptr) {
        std::vector<uint64_t> host = build_fixed_table_host();
        ZKW_TRY(ensure_buffer(ctx, t, host.size() * 8));
        ZKW_CUDA(ctx, cudaMemcpyAsync(t.ptr, host.data(), host.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        ZKW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    *out_dev = (const uint64_t*)t.ptr;
    return ZKW_OK;
}

int fixed_base_mul_dev(zkw_ctx* ctx, const uint64_t* scalars_dev, size_t n, uint64_t* out_xy_dev) {
    const uint64_t* table = nullptr;
    ZKW_TRY(fixed_base_table(ctx, &table));
    { ProfScope ps_(ctx, "fixed_base_kernel"); fixed_base_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>((const uint4*)scalars_dev, (const uint4*)table, (uint4*)out_xy_dev, n); }
    ZKW_LAUNCHED(ctx);
    return ZKW_OK;
}

int srs_setup(zkw_ctx* ctx, unsigned k, const uint64_t tau_m[4]) {
    if (k > 26) return ZKW_ERR_INVALID;
    const size_t n = (size_t)1 << k;
    Fr tau;
    memcpy(tau.l, tau_m, 32);
    DomainConsts dc;
    domain_consts(k, k, &dc);
    Fr omega;
    memcpy(omega.l, dc.omega, 32);
    Fr n_inv;
    memcpy(n_inv.l, dc.n_inv, 32);
    Fr tn = tau;
    for (unsigned i = 0; i < k; i++) tn = tn.sqr();
    Fr c = (tn - Fr::one()) * n_inv;
    if (c.is_zero()) return ZKW_ERR_INVALID;  // tau in the evaluation domain: degenerate SRS

    DeviceBuffer sc, pw;
    ZKW_TRY(ensure_buffer(ctx, sc, n * 32));
    ZKW_TRY(ensure_buffer(ctx, pw, n * 32));
    for (int which = 0; which < 2; which++) {
        MsmBasis& b = ctx->bases[which];
        msm_free_basis(b);
        ZKW_CUDA(ctx, cudaMalloc((void**)&b.points, n * 64));
        b.n = n;
    }
    const unsigned run = 64;
    const unsigned threads = (unsigned)((n + run - 1) / run);
    // g[i] = tau^i G
    { ProfScope ps_(ctx, "powers_kernel"); powers_kernel<<<(threads + 127) / 128, 128, 0, ctx->stream>>>((uint4*)sc.ptr, tau, n, run); }
    ZKW_LAUNCHED(ctx);
    ZKW_TRY(fixed_base_mul_dev(ctx, (const uint64_t*)sc.ptr, n, ctx->bases[ZKW_BASES_G].points));
    // g_lagrange[i] = L_i(tau) G
    { ProfScope ps_(ctx, "powers_kernel"); powers_kernel<<<(threads + 127) / 128, 128, 0, ctx->stream>>>((uint4*)pw.ptr, omega, n, run); }
    ZKW_LAUNCHED(ctx);
    { ProfScope ps_(ctx, "lagrange_scalars_kernel"); lagrange_scalars_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>((const uint4*)pw.ptr, (uint4*)sc.ptr, tau, c, n); }
    ZKW_LAUNCHED(ctx);
    ZKW_TRY(fixed_base_mul_dev(ctx, (const uint64_t*)sc.ptr, n, ctx->bases[ZKW_BASES_G_LAGRANGE].points));
    ZKW_TRY(msm_prepare_basis(ctx, ctx->bases[ZKW_BASES_G]));
    ZKW_TRY(msm_prepare_basis(ctx, ctx->bases[ZKW_BASES_G_LAGRANGE]));
    ZKW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(sc.ptr);
    cudaFree(pw.ptr);
    return ZKW_OK;
}

}  // namespace zkw
