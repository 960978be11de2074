// ntt.cu — radix-2 NTT over BN254 Fr for sm_100a: the device replacement of
// halo2_proofs::arithmetic::best_fft and of the EvaluationDomain transforms built on it
// (reached from the reference through create_proof / keygen, halo2-circuits/src/ecc/ecdsa_p256.rs:259-260,
// 366-373, 416-423, 555-562).  Contract: natural order in and out, out[i] = sum_j a[j] omega^(ij).
//
// Layout of the computation.  The log_n butterfly stages (decimation in time) are cut into passes
// of B <= 7 stages.  One CTA owns a tile of 2^10 elements = 2^B consecutive butterfly positions
// ("mid") x 2^(10-B) independent columns, staged in 32 KB of shared memory as two uint4 planes
// (low / high 16 bytes of each element) so that a warp's 128-bit accesses are conflict-light.
// Inside a pass each thread keeps 8 elements in registers and runs up to three butterfly stages
// on them before the tile is exchanged through shared memory again.  The bit-reversal permutation
// is folded into the first pass's gather (one 32-byte sector per element), the coset pre-scaling
// zeta^(i mod 3) of coeff_to_extended into the same gather, and the 1/n (and zeta^-(i mod 3))
// post-scaling of the inverse transforms into the last pass's store, so every pass reads and writes
// each element exactly once: algorithmic traffic 64 B per element per pass.
// Twiddles omega^i (i < n/2) come from a per-(omega, log_n) table cached in the context.
//
// The kernel is integer-ALU bound (one 254-bit Montgomery product per butterfly), not HBM bound;
// see DESIGN.md for the roofline.
#include "common.cuh"

namespace zkw {

constexpr int kTileLog = 10;
constexpr int kNttThreads = 128;
constexpr int kMaxPassBits = 7;
#ifndef ZKW_NTT_MIN_BLOCKS
#define ZKW_NTT_MIN_BLOCKS 4
#endif

struct NttPassArgs {
    const uint4* src;
    uint4* dst;
    const uint4* tw;
    int log_n;       // transform size
    int src_log_n;   // first pass: source has 2^src_log_n elements, the rest read as zero
    int s0;          // stages already done
    int B;           // stages in this pass
    int tl;          // log2 of the tile (min(kTileLog, log_n))
    int first;       // gather with bit-reversed index
    int coset;       // first pass: multiply source element j by zeta^(j mod 3)
    int scale;       // last pass: multiply output element i by scale3[i mod 3]
    Fr zeta, zeta2;
    Fr scale3[3];
};

__device__ __forceinline__ Fr lds_fr(const uint4* lo, const uint4* hi, int p) {
    uint4 a = lo[p], b = hi[p];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void sts_fr(uint4* lo, uint4* hi, int p, const Fr& v) {
    lo[p] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    hi[p] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

// R butterfly stages on 2^R register-resident elements per task.
template <int R>
__device__ __forceinline__ void ntt_round(uint4* slo, uint4* shi, const NttPassArgs& a, unsigned tile, int t) {
    const int C = a.tl - a.B;
    const int ntasks = 1 << (a.tl - R);
    const int s = a.s0 + t;
    for (int task = threadIdx.x; task < ntasks; task += kNttThreads) {
        const int cc = task & ((1 << C) - 1);
        const int g = task >> C;
        const unsigned c = (tile << C) + cc;
        const unsigned lo = c & ((1u << a.s0) - 1u);
        const int glow = g & ((1 << t) - 1);
        const int mid_base = ((g >> t) << (t + R)) | glow;
        const unsigned jm = ((unsigned)glow << a.s0) | lo;
        Fr x[1 << R];
#pragma unroll
        for (int e = 0; e < (1 << R); e++) x[e] = lds_fr(slo, shi, ((mid_base | (e << t)) << C) + cc);
#pragma unroll
        for (int u = 0; u < R; u++) {
            const int sh = a.log_n - (s + u + 1);
#pragma unroll
            for (int e = 0; e < (1 << R); e++) {
                if (e & (1 << u)) continue;
                const unsigned el = e & ((1 << u) - 1);
                Fr tv;
                if (s + u == 0) {
                    tv = x[e | (1 << u)];  // stage 0: every twiddle is 1
                } else {
#ifdef ZKW_NTT_FAKE_TW   // timing experiment only (wrong results): every twiddle load hits L1
                    const unsigned idx = ((jm + (el << s)) << sh) & 63u;
#else
                    const unsigned idx = (jm + (el << s)) << sh;
#endif
                    Fr w = Fr::load_nc(a.tw + 2 * (size_t)idx);
                    tv = x[e | (1 << u)] * w;
                }
                Fr uu = x[e];
                x[e] = uu + tv;
                x[e | (1 << u)] = uu - tv;
            }
        }
#pragma unroll
        for (int e = 0; e < (1 << R); e++) sts_fr(slo, shi, ((mid_base | (e << t)) << C) + cc, x[e]);
    }
}

__global__ void __launch_bounds__(kNttThreads, ZKW_NTT_MIN_BLOCKS) ntt_pass_kernel(const NttPassArgs a) {
    __shared__ uint4 slo[1 << kTileLog];
    __shared__ uint4 shi[1 << kTileLog];
    const unsigned tile = blockIdx.x;
    const int C = a.tl - a.B;
    const int tsize = 1 << a.tl;
    // ---- gather the tile ----
    for (int q = threadIdx.x; q < tsize; q += kNttThreads) {
        int mid, cc;
        if (a.s0 == 0) { mid = q & ((1 << a.B) - 1); cc = q >> a.B; }
        else { cc = q & ((1 << C) - 1); mid = q >> C; }
        const unsigned c = (tile << C) + cc;
        const size_t i = ((size_t)(c >> a.s0) << (a.s0 + a.B)) | ((size_t)mid << a.s0) | (c & ((1u << a.s0) - 1u));
        Fr v;
        if (a.first) {
            const unsigned j = __brev((unsigned)i) >> (32 - a.log_n);
            if ((j >> a.src_log_n) != 0) {
                v = Fr::zero();
            } else {
                v = Fr::load(a.src + 2 * (size_t)j);
                if (a.coset) {
                    const unsigned m3 = j % 3u;
                    if (m3 == 1) v = v * a.zeta;
                    else if (m3 == 2) v = v * a.zeta2;
                }
            }
        } else {
            v = Fr::load(a.src + 2 * i);
        }
        sts_fr(slo, shi, (mid << C) + cc, v);
    }
    __syncthreads();
    // ---- butterfly rounds ----
    int t = 0;
    while (t < a.B) {
        const int r = (a.B - t >= 3) ? 3 : (a.B - t);
        if (r == 3) ntt_round<3>(slo, shi, a, tile, t);
        else if (r == 2) ntt_round<2>(slo, shi, a, tile, t);
        else ntt_round<1>(slo, shi, a, tile, t);
        t += r;
        __syncthreads();
    }
    // ---- scatter the tile ----
    for (int q = threadIdx.x; q < tsize; q += kNttThreads) {
        int mid, cc;
        if (a.s0 == 0) { mid = q & ((1 << a.B) - 1); cc = q >> a.B; }
        else { cc = q & ((1 << C) - 1); mid = q >> C; }
        const unsigned c = (tile << C) + cc;
        const size_t i = ((size_t)(c >> a.s0) << (a.s0 + a.B)) | ((size_t)mid << a.s0) | (c & ((1u << a.s0) - 1u));
        Fr v = lds_fr(slo, shi, (mid << C) + cc);
        if (a.scale) v = v * a.scale3[i % 3];
        v.store(a.dst + 2 * i);
    }
}

// tw[i] = omega^i for i < count: each thread seeds its run with a square-and-multiply power.
__global__ void twiddle_kernel(uint4* tw, Fr omega, unsigned count, unsigned run) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned start = t * run;
    if (start >= count) return;
    Fr w = omega.pow((uint64_t)start);
    const unsigned end = min(count, start + run);
    for (unsigned i = start; i < end; i++) {
        w.store(tw + 2 * (size_t)i);
        w = w * omega;
    }
}

// n == 1 or scaling-only helper: dst[i] = src[i] * scale3[i mod 3]
__global__ void scale_kernel(const uint4* src, uint4* dst, size_t n, Fr s0, Fr s1, Fr s2, int has_scale) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr v = Fr::load(src + 2 * i);
    if (has_scale) {
        unsigned m = (unsigned)(i % 3);
        v = v * (m == 0 ? s0 : (m == 1 ? s1 : s2));
    }
    v.store(dst + 2 * i);
}

static Fr fr_from_host(const uint64_t v[4]) {
    Fr r;
    memcpy(r.l, v, 32);
    return r;
}

int ntt_get_twiddles(zkw_ctx* ctx, const uint64_t omega[4], unsigned log_n, const uint64_t** out_dev) {
    TwiddleKey key;
    memcpy(key.omega.data(), omega, 32);
    key.log_n = log_n;
    auto it = ctx->twiddles.find(key);
    if (it != ctx->twiddles.end()) {
        *out_dev = (const uint64_t*)it->second.ptr;
        return ZKW_OK;
    }
    const unsigned count = log_n == 0 ? 1u : (1u << (log_n - 1));
    DeviceBuffer buf;
    ZKW_TRY(ensure_buffer(ctx, buf, (size_t)count * 32));
    const unsigned run = 64;
    const unsigned threads = (count + run - 1) / run;
    { ProfScope ps_(ctx, "twiddle_kernel"); twiddle_kernel<<<(threads + 127) / 128, 128, 0, ctx->stream>>>((uint4*)buf.ptr, fr_from_host(omega), count, run); }
    ZKW_LAUNCHED(ctx);
    ctx->twiddles[key] = buf;
    *out_dev = (const uint64_t*)buf.ptr;
    return ZKW_OK;
}

int ntt_run(zkw_ctx* ctx, const uint64_t* src_dev, unsigned src_log_n, uint64_t* dst_dev, unsigned log_n,
            const uint64_t omega[4], bool coset_in, const uint64_t* scale3, cudaStream_t stream) {
    // work on the auxiliary stream gets its own scratch so that it may overlap main-stream transforms
    const bool aux = stream && stream != ctx->stream;
    cudaStream_t st = stream ? stream : ctx->stream;
    DeviceBuffer& scratch = aux ? ctx->ntt_scratch_aux : ctx->ntt_scratch;
    if (log_n > 28 || src_log_n > log_n) return ZKW_ERR_INVALID;
    const size_t n = (size_t)1 << log_n;
    if (log_n == 0) {
        Fr s0 = Fr::one(), s1 = s0, s2 = s0;
        if (scale3) { s0 = fr_from_host(scale3); s1 = fr_from_host(scale3 + 4); s2 = fr_from_host(scale3 + 8); }
        { ProfScope ps_(ctx, "scale_kernel", st); scale_kernel<<<1, 32, 0, st>>>((const uint4*)src_dev, (uint4*)dst_dev, 1, s0, s1, s2, scale3 != nullptr); }
        ZKW_LAUNCHED(ctx);
        return ZKW_OK;
    }
    const uint64_t* tw = nullptr;
    ZKW_TRY(ntt_get_twiddles(ctx, omega, log_n, &tw));
    const int npass = (int)((log_n + kMaxPassBits - 1) / kMaxPassBits);
    const int base = (int)log_n / npass, rem = (int)log_n % npass;
    // The first pass permutes (bit reversal), so it cannot run in place when there are several
    // tiles: route it through the scratch buffer unless src and dst already differ.
    const bool in_place = (const void*)src_dev == (const void*)dst_dev;
    uint64_t* tmp = nullptr;
    if (in_place && log_n > (unsigned)kTileLog) {
        ZKW_TRY(ensure_buffer(ctx, scratch, n * 32));
        tmp = (uint64_t*)scratch.ptr;
    }
    // single-tile transforms gather the whole input before the first barrier, so in place is safe
    NttPassArgs a;
    memset(&a, 0, sizeof(a));
    a.tw = (const uint4*)tw;
    a.log_n = (int)log_n;
    a.src_log_n = (int)src_log_n;
    a.tl = (int)(log_n < (unsigned)kTileLog ? log_n : kTileLog);
    a.coset = coset_in ? 1 : 0;
    {
        // zeta = 7^((r-1)/3) in Montgomery form
        static const uint64_t zeta_m[4] = {0x93e7cede4a0329b3ULL, 0x7d4fdca77a96c167ULL, 0x8be4ba08b19a750aULL, 0x1cbd5653a5661c25ULL};
        a.zeta = fr_from_host(zeta_m);
        a.zeta2 = a.zeta * a.zeta;
    }
    if (scale3) {
        a.scale3[0] = fr_from_host(scale3);
        a.scale3[1] = fr_from_host(scale3 + 4);
        a.scale3[2] = fr_from_host(scale3 + 8);
    }
    int s0 = 0;
    const uint64_t* cur_src = src_dev;
    for (int p = 0; p < npass; p++) {
        const int B = base + (p < rem ? 1 : 0);
        a.s0 = s0;
        a.B = B;
        a.first = (p == 0);
        a.scale = (p == npass - 1 && scale3) ? 1 : 0;
        uint64_t* out = dst_dev;
        if (p == 0 && tmp && npass > 1) out = tmp;             // a -> tmp, later passes tmp -> ... -> a
        if (p > 0 && p < npass - 1 && tmp) out = tmp;          // middle passes stay in tmp (tile-local in place)
        if (p == 0 && tmp && npass == 1) out = tmp;            // cannot happen (npass==1 => log_n<=7), kept for clarity
        a.src = (const uint4*)cur_src;
        a.dst = (uint4*)out;
        const unsigned tiles = (unsigned)(n >> a.tl);
        { ProfScope ps_(ctx, "ntt_pass_kernel", st); ntt_pass_kernel<<<tiles, kNttThreads, 0, st>>>(a); }
        ZKW_LAUNCHED(ctx);
        cur_src = out;
        s0 += B;
    }
    if ((const void*)cur_src != (const void*)dst_dev) {
        ZKW_CUDA(ctx, cudaMemcpyAsync(dst_dev, cur_src, n * 32, cudaMemcpyDeviceToDevice, st));
    }
    return ZKW_OK;
}

}  // namespace zkw
