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
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.cuh"

namespace zkw {

#ifndef ZKW_NTT_TILE_LOG
#define ZKW_NTT_TILE_LOG 10
#endif
#ifndef ZKW_NTT_THREADS
#define ZKW_NTT_THREADS 128
#endif
#ifndef ZKW_NTT_MAX_PASS_BITS
#define ZKW_NTT_MAX_PASS_BITS 7
#endif
#ifndef ZKW_NTT_MIN_BLOCKS
#define ZKW_NTT_MIN_BLOCKS 4
#endif
constexpr int kTileLog = ZKW_NTT_TILE_LOG;
constexpr int kNttThreads = ZKW_NTT_THREADS;
constexpr int kMaxPassBits = ZKW_NTT_MAX_PASS_BITS;
// two uint4 planes of one tile; with twiddle staging also the last stage's twiddles (half a tile of elements) and
// one mbarrier
constexpr size_t kNttTileBytes = (size_t)2 * sizeof(uint4) << kTileLog;
constexpr size_t kNttSmemBytes = ((size_t)3 * sizeof(uint4) << kTileLog) + 16;

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
    int last;        // last pass: outputs leave the lazy range [0, 4r) for [0, r)
    int zero_stages; // first pass of a zero-padded transform: in stages below this the odd input of every butterfly is zero
    Fr zeta, zeta2;
    Fr scale3[3];
    const uint4* stw;   // staged twiddles of this pass's LAST stage, one contiguous block per tile group, or nullptr
    unsigned stw_group_mask;  // tile -> group = tile & mask
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
// tlo / thi: the staged twiddles of the pass's last stage (entry (j << C) + cc, j = butterfly index inside the
// stage), or nullptr when this round does not contain that stage / nothing is staged.
template <int R>
__device__ __forceinline__ void ntt_round(uint4* slo, uint4* shi, const NttPassArgs& a, unsigned tile, int t,
                                          const uint4* tlo, const uint4* thi) {
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
                // elements live in [0, 4r) between stages (Harvey): one conditional subtraction per butterfly
                Fr tv;
                if (s + u < a.zero_stages) {
                    // zero-padded source (coeff_to_extended: 2^k coefficients in a 2^(k+2) transform): after the bit-reversed
                    // gather the elements whose low log_n - src_log_n index bits are not all zero are zero, so in the first
                    // log_n - src_log_n stages every butterfly is (x, 0) -> (x, x): no twiddle, no product
                    x[e | (1 << u)] = x[e];
                    continue;
                }
                if (s + u == 0) {
                    tv = x[e | (1 << u)].reduced_2m();  // stage 0: every twiddle is 1
                } else {
#ifdef ZKW_NTT_FAKE_TW   // timing experiment only (wrong results): every twiddle load hits L1
                    const unsigned idx = ((jm + (el << s)) << sh) & 63u;
#else
                    const unsigned idx = (jm + (el << s)) << sh;
#endif
                    Fr w;
                    if (u == R - 1 && tlo) w = lds_fr(tlo, thi, (int)((((unsigned)glow | (el << t)) << C) + cc));
                    else w = Fr::load_nc(a.tw + 2 * (size_t)idx);
                    tv = Fr::mul_lazy(x[e | (1 << u)], w);   // < 2r for an input below 4r and a canonical twiddle
                }
                const Fr uu = x[e].reduced_2m();
                x[e] = Fr::add_raw(uu, tv);                      // < 4r
                x[e | (1 << u)] = Fr::sub_plus_2m(uu, tv);       // < 4r
            }
        }
#pragma unroll
        for (int e = 0; e < (1 << R); e++) sts_fr(slo, shi, ((mid_base | (e << t)) << C) + cc, x[e]);
    }
}

__global__ void __launch_bounds__(kNttThreads, ZKW_NTT_MIN_BLOCKS) ntt_pass_kernel(const NttPassArgs a) {
    extern __shared__ uint4 ntt_smem[];
    uint4* slo = ntt_smem;
    uint4* shi = ntt_smem + (1 << kTileLog);
    uint4* tlo = ntt_smem + (2 << kTileLog);
    uint64_t* bar = reinterpret_cast<uint64_t*>(ntt_smem + (3 << kTileLog));
    const unsigned tile = blockIdx.x;
    const int C = a.tl - a.B;
    const int tsize = 1 << a.tl;
    // ---- stage the last stage's twiddles: one bulk copy, in flight while the tile is gathered and the
    // earlier rounds run ----
    const int stw_entries = 1 << (a.tl - 1);   // 2^(B-1) butterflies x 2^C columns
    const uint4* thi = tlo + stw_entries;
    if (a.stw && threadIdx.x == 0) {
        mbar_init(bar, 1);
        bulk_load(tlo, a.stw + 2 * (size_t)(tile & a.stw_group_mask) * stw_entries, (uint32_t)stw_entries * 32u, bar);
    }
    // ---- gather the tile ----
    // Full tiles (the only case above 2^10 elements) issue all of a thread's loads before touching any of them: 16 128-bit
    // loads in flight per thread instead of one element at a time (the bit-reversed first pass reads isolated 32-byte
    // sectors, each a DRAM round trip).
    auto tile_index = [&](int q, int& mid, int& cc) -> size_t {
        if (a.s0 == 0) { mid = q & ((1 << a.B) - 1); cc = q >> a.B; }
        else { cc = q & ((1 << C) - 1); mid = q >> C; }
        const unsigned c = (tile << C) + cc;
        return ((size_t)(c >> a.s0) << (a.s0 + a.B)) | ((size_t)mid << a.s0) | (c & ((1u << a.s0) - 1u));
    };
    if (tsize == (kNttThreads << 3)) {
        Fr v[8];
        unsigned src_j[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
            int mid, cc;
            const size_t i = tile_index(threadIdx.x + e * kNttThreads, mid, cc);
            if (a.first) {
                const unsigned j = __brev((unsigned)i) >> (32 - a.log_n);
                src_j[e] = j;
                v[e] = (j >> a.src_log_n) != 0 ? Fr::zero() : Fr::load(a.src + 2 * (size_t)j);
            } else {
                src_j[e] = 0;
                v[e] = Fr::load(a.src + 2 * i);
            }
        }
#pragma unroll
        for (int e = 0; e < 8; e++) {
            int mid, cc;
            tile_index(threadIdx.x + e * kNttThreads, mid, cc);
            if (a.first && a.coset && (src_j[e] >> a.src_log_n) == 0) {
                const unsigned m3 = src_j[e] % 3u;
                if (m3 == 1) v[e] = v[e] * a.zeta;
                else if (m3 == 2) v[e] = v[e] * a.zeta2;
            }
            sts_fr(slo, shi, (mid << C) + cc, v[e]);
        }
    } else {
    for (int q = threadIdx.x; q < tsize; q += kNttThreads) {
        int mid, cc;
        const size_t i = tile_index(q, mid, cc);
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
    }
    __syncthreads();
    // ---- butterfly rounds ----
    int t = 0;
    while (t < a.B) {
        const int r = (a.B - t >= 3) ? 3 : (a.B - t);
        const bool last = t + r == a.B;
        if (last && a.stw) mbar_wait(bar, 0);   // the barrier was initialised before the __syncthreads above
        const uint4* wl = (last && a.stw) ? tlo : nullptr;
        if (r == 3) ntt_round<3>(slo, shi, a, tile, t, wl, thi);
        else if (r == 2) ntt_round<2>(slo, shi, a, tile, t, wl, thi);
        else ntt_round<1>(slo, shi, a, tile, t, wl, thi);
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
        if (a.last) {        // leave the lazy range: canonical output
            if (a.scale) v = v * a.scale3[i % 3];               // Montgomery product of v < 4r with a canonical scale: < r
            else v = v.normalized_4m();
        }
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

// Compact copy of the twiddles a pass's last stage needs, laid out the way a tile consumes them: for tile
// group q (the tile's low column bits), the block [q] holds plane 0 (low 16 bytes) of entries (j << C) + cc,
// then plane 1 - so that ONE contiguous bulk copy stages a tile's 2^(tl-1) twiddles into shared memory.
// entry (q, j, cc) = omega^(((j << s0) | (q << C) | cc) << (log_n - s0 - B)).
__global__ void stage_twiddle_kernel(const uint4* __restrict__ tw, uint4* __restrict__ out, int log_n, int s0, int B, int C, unsigned groups) {
    const unsigned per = 1u << (B - 1 + C);
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)groups * per) return;
    const unsigned q = (unsigned)(i / per), e = (unsigned)(i % per);
    const unsigned cc = e & ((1u << C) - 1u), j = e >> C;
    const size_t idx = (size_t)(((j << s0) | (q << C) | cc)) << (log_n - s0 - B);
    out[2 * (size_t)q * per + e] = tw[2 * idx];
    out[2 * (size_t)q * per + per + e] = tw[2 * idx + 1];
}

static int ntt_get_staged(zkw_ctx* ctx, const uint64_t omega[4], const uint64_t* tw, int log_n, int s0, int B, int C,
                          const uint64_t** out_dev, unsigned* group_mask) {
    *out_dev = nullptr;
    *group_mask = 0;
    if (s0 < C || s0 == 0 || B < 1) return ZKW_OK;         // first pass: 2^(B-1) twiddles shared by every tile, L1 keeps them
    const unsigned groups = 1u << (s0 - C);
    const size_t bytes = (size_t)groups * ((size_t)32 << (B - 1 + C));
    if (bytes > ((size_t)1 << 30)) return ZKW_OK;          // huge transforms: plain loads
    StagedTwiddleKey key;
    memcpy(key.omega.data(), omega, 32);
    key.log_n = (unsigned)log_n; key.s0 = (unsigned)s0; key.B = (unsigned)B;
    auto it = ctx->staged_twiddles.find(key);
    if (it == ctx->staged_twiddles.end()) {
        DeviceBuffer buf;
        ZKW_TRY(ensure_buffer(ctx, buf, bytes));
        const size_t total = (size_t)groups << (B - 1 + C);
        // built on the context's main stream, like the main table; a transform on another stream that needs it
        // right away is ordered by the synchronisation below (first use only)
        { ProfScope ps_(ctx, "stage_twiddle_kernel"); stage_twiddle_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((const uint4*)tw, (uint4*)buf.ptr, log_n, s0, B, C, groups); }
        ZKW_LAUNCHED(ctx);
        ZKW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        it = ctx->staged_twiddles.emplace(key, buf).first;
    }
    *out_dev = (const uint64_t*)it->second.ptr;
    *group_mask = groups - 1;
    return ZKW_OK;
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
    // cold path (once per (omega, log n) and context): the table is built on the main stream but read by whichever
    // stream the transform runs on (aux stream, MSM lanes' callers), so finish it before anyone can see the pointer
    ZKW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->twiddles[key] = buf;
    *out_dev = (const uint64_t*)buf.ptr;
    return ZKW_OK;
}

int ntt_run(zkw_ctx* ctx, const uint64_t* src_dev, unsigned src_log_n, uint64_t* dst_dev, unsigned log_n,
            const uint64_t omega[4], bool coset_in, const uint64_t* scale3, cudaStream_t stream) {
    // work on the auxiliary stream gets its own scratch so that it may overlap main-stream transforms
    const bool aux = stream && stream != ctx->stream;
    cudaStream_t st = stream ? stream : ctx->stream;
    int aux_i = 0;
    for (int i = 1; i < zkw_ctx::kAuxStreams; i++) if (stream && stream == ctx->aux_stream[i]) aux_i = i;
    DeviceBuffer& scratch = aux ? ctx->ntt_scratch_aux[aux_i] : ctx->ntt_scratch;
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
    // Twiddle staging by TMA is OFF unless ZKW_NTT_STAGE_TWIDDLES=1.  Measured (profiles/r1c_ntt_tuning.md): against
    // the same kernel without staging it gains 1.5-2 %, but its 16 KB per CTA lift four resident CTAs from 132 KB to
    // 196+ KB of shared memory, the SM's carve-out goes to 228 KB, and the L1 left over for this kernel's early-stage
    // twiddles and for the MSM kernels running next to it shrinks from 96 KB to nothing: the k = 19 proof takes
    // 27.6-27.8 ms with the 48 KB kernel (staged or not) against 27.2-27.3 ms with the 32 KB one.
    const char* stage_env = getenv("ZKW_NTT_STAGE_TWIDDLES");
    const bool stage = stage_env && stage_env[0] == '1';
    const size_t smem_bytes = stage ? kNttSmemBytes : kNttTileBytes;
    if (smem_bytes > 48 * 1024 && !ctx->ntt_attr_set) {
        ZKW_CUDA(ctx, cudaFuncSetAttribute(ntt_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNttSmemBytes));
        ctx->ntt_attr_set = true;
    }
    // stages per pass: as even as possible with at most kMaxPassBits each; ZKW_NTT_PLAN_<log_n>="9,6,6" overrides
    // the split (tuning aid: each entry <= the tile's log size, entries sum to log_n)
    std::vector<int> plan;
    {
        const int np = (int)((log_n + kMaxPassBits - 1) / kMaxPassBits);
        const int base = (int)log_n / np, rem = (int)log_n % np;
        for (int p = 0; p < np; p++) plan.push_back(base + (p < rem ? 1 : 0));
        char name[32];
        snprintf(name, sizeof(name), "ZKW_NTT_PLAN_%u", log_n);
        if (const char* env = getenv(name)) {
            std::vector<int> alt;
            int sum = 0;
            bool ok = true;
            for (const char* q = env; *q;) {
                char* end = nullptr;
                const long v = strtol(q, &end, 10);
                if (end == q || v < 1 || v > (long)(log_n < (unsigned)kTileLog ? log_n : kTileLog)) { ok = false; break; }
                alt.push_back((int)v);
                sum += (int)v;
                q = *end == ',' ? end + 1 : end;
            }
            if (ok && sum == (int)log_n) plan = alt;
        }
    }
    const int npass = (int)plan.size();
    static const bool zero_skip = !getenv("ZKW_NTT_NO_ZERO_SKIP");   // A/B knob
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
        const int B = plan[p];
        a.s0 = s0;
        a.B = B;
        a.first = (p == 0);
        a.scale = (p == npass - 1 && scale3) ? 1 : 0;
        a.last = (p == npass - 1) ? 1 : 0;
        a.zero_stages = (p == 0 && zero_skip) ? std::min<int>((int)(log_n - src_log_n), B) : 0;
        uint64_t* out = dst_dev;
        if (p == 0 && tmp && npass > 1) out = tmp;             // a -> tmp, later passes tmp -> ... -> a
        if (p > 0 && p < npass - 1 && tmp) out = tmp;          // middle passes stay in tmp (tile-local in place)
        if (p == 0 && tmp && npass == 1) out = tmp;            // cannot happen (npass==1 => log_n<=7), kept for clarity
        a.src = (const uint4*)cur_src;
        a.dst = (uint4*)out;
        a.stw = nullptr;
        a.stw_group_mask = 0;
        if (stage && a.tl == kTileLog) {
            const uint64_t* stw = nullptr;
            ZKW_TRY(ntt_get_staged(ctx, omega, tw, (int)log_n, s0, B, a.tl - B, &stw, &a.stw_group_mask));
            a.stw = (const uint4*)stw;
        }
        const unsigned tiles = (unsigned)(n >> a.tl);
        { ProfScope ps_(ctx, "ntt_pass_kernel", st); ntt_pass_kernel<<<tiles, kNttThreads, smem_bytes, st>>>(a); }
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
