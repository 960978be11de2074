// capi.cu — the extern "C" surface declared in include/zkw_b200.h: context, SRS residency, and the
// host-pointer / device-pointer entry points a patched halo2_proofs binds in place of best_multiexp,
// best_fft, EvaluationDomain::{lagrange_to_coeff, coeff_to_extended, extended_to_coeff} and
// evaluate_h (call sites in the reference: halo2-circuits/src/ecc/ecdsa_p256.rs:259-260, 366-373,
// 416-423, 555-562).  No CPU fallback lives here: every path ends in a kernel launch or an error.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include "common.cuh"

namespace zkw {

int set_cuda_error(zkw_ctx* ctx, cudaError_t e, const char* what) {
    if (ctx) {
        ctx->last_err = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    }
    cudaGetLastError();  // clear the sticky-less error state
    if (e == cudaErrorMemoryAllocation) return ZKW_ERR_OOM;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice) return ZKW_ERR_NO_DEVICE;
    return ZKW_ERR_CUDA;
}

int ensure_buffer(zkw_ctx* ctx, DeviceBuffer& b, size_t bytes) {
    if (b.bytes >= bytes && b.ptr) return ZKW_OK;
    if (b.ptr) {
        // the old area may still be in use by work queued on ANY of the context's streams (main, aux, MSM lanes)
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) return set_cuda_error(ctx, e, "cudaDeviceSynchronize");
        cudaFree(b.ptr);
        b.ptr = nullptr;
        b.bytes = 0;
    }
    size_t want = bytes < 256 ? 256 : bytes;
    cudaError_t e = cudaMalloc(&b.ptr, want);
    if (e != cudaSuccess) { b.ptr = nullptr; return set_cuda_error(ctx, e, "cudaMalloc"); }
    b.bytes = want;
    return ZKW_OK;
}

// ---- host-side domain constants (EvaluationDomain::new restated with the device field class) ----
static Fr fr_pow_host(Fr b, uint64_t e) { return b.pow(e); }

void domain_consts(unsigned k, unsigned ext_k, DomainConsts* out) {
    static std::mutex mu;
    static std::map<std::pair<unsigned, unsigned>, DomainConsts> cache;
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_pair(k, ext_k);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return; }
    // ROOT_OF_UNITY = 7^((r-1)/2^28), ZETA = 7^((r-1)/3), Montgomery form
    static const uint64_t root_m[4] = {0x9632c7c5b639feb8ULL, 0x985ce3400d0ff299ULL, 0xb2dd880001b0ecd8ULL, 0x1d69070d6d98ce29ULL};
    static const uint64_t zeta_m[4] = {0x93e7cede4a0329b3ULL, 0x7d4fdca77a96c167ULL, 0x8be4ba08b19a750aULL, 0x1cbd5653a5661c25ULL};
    DomainConsts d;
    memset(&d, 0, sizeof(d));
    d.k = k; d.ext_k = ext_k;
    Fr w; memcpy(w.l, root_m, 32);
    for (unsigned i = ext_k; i < 28; i++) w = w.sqr();
    Fr ext_omega = w;
    for (unsigned i = k; i < ext_k; i++) w = w.sqr();
    Fr omega = w;
    Fr zeta; memcpy(zeta.l, zeta_m, 32);
    Fr zeta_inv = zeta.sqr();
    Fr two = Fr::one() + Fr::one();
    Fr n_inv = fr_pow_host(two, k).inv(), en_inv = fr_pow_host(two, ext_k).inv();
    Fr omega_inv = omega.inv(), ext_omega_inv = ext_omega.inv();
    memcpy(d.omega, omega.l, 32); memcpy(d.omega_inv, omega_inv.l, 32);
    memcpy(d.ext_omega, ext_omega.l, 32); memcpy(d.ext_omega_inv, ext_omega_inv.l, 32);
    memcpy(d.zeta, zeta.l, 32); memcpy(d.zeta_inv, zeta_inv.l, 32);
    memcpy(d.n_inv, n_inv.l, 32); memcpy(d.ext_n_inv, en_inv.l, 32);
    Fr s1 = en_inv * zeta_inv, s2 = en_inv * zeta;  // zeta^-1 = zeta^2, zeta^-2 = zeta
    memcpy(d.ext_scale3, en_inv.l, 32); memcpy(d.ext_scale3 + 4, s1.l, 32); memcpy(d.ext_scale3 + 8, s2.l, 32);
    for (int i = 0; i < 3; i++) memcpy(d.n_scale3 + 4 * i, n_inv.l, 32);
    unsigned m = 1u << (ext_k - k);
    Fr cur = zeta;
    for (unsigned i = 0; i < m && i < 16; i++) {
        Fr t = fr_pow_host(cur, 1ull << k) - Fr::one();
        Fr ti = t.inv();
        memcpy(d.t_evals[i], ti.l, 32);
        cur = cur * ext_omega;
    }
    cache[key] = d;
    *out = d;
}

int stream_priority(int level) {
    int least = 0, greatest = 0;
    if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) { cudaGetLastError(); return 0; }
    // numerically lower = more urgent; level 0 = background, 2 = most urgent
    if (level <= 0) return least;
    if (level >= 2) return greatest;
    return (least + greatest) / 2;
}

// ---- per-kernel timing ---------------------------------------------------------------------------
static cudaEvent_t take_event(zkw_ctx* ctx) {
    if (!ctx->event_pool.empty()) {
        cudaEvent_t e = ctx->event_pool.back();
        ctx->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return e;
}

ProfScope::ProfScope(zkw_ctx* c, const char* name, cudaStream_t s) : ctx(c), stream(s ? s : c->stream) {
    if (!c->profiling || (!c->prof_filter.empty() && c->prof_filter != name)) return;
    cudaEvent_t start = take_event(c);
    stop = take_event(c);
    if (!start || !stop) { stop = nullptr; return; }
    cudaEventRecord(start, stream);
    c->prof_pending.push_back({name, start, stop, stream});
}
ProfScope::~ProfScope() {
    if (stop) cudaEventRecord(stop, stream);
}

static void profile_collect(zkw_ctx* ctx) {
    if (ctx->prof_pending.empty()) return;
    cudaStreamSynchronize(ctx->stream);
    for (int i = 1; i < zkw_ctx::kMsmLanes; i++) if (ctx->lane_stream[i]) cudaStreamSynchronize(ctx->lane_stream[i]);
    for (int i = 0; i < zkw_ctx::kAuxStreams; i++) if (ctx->aux_stream[i]) cudaStreamSynchronize(ctx->aux_stream[i]);
    // ZKW_TIMELINE=<file>: also append one "name,stream,start_ms,duration_ms" line per launch (start relative
    // to the first launch profiled), which is enough to see what overlaps what without a system profiler
    const char* tl_path = getenv("ZKW_TIMELINE");
    FILE* tl = tl_path ? fopen(tl_path, "a") : nullptr;
    if (tl && !ctx->prof_ref) ctx->prof_ref = ctx->prof_pending.front().start;
    for (auto& r : ctx->prof_pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.start, r.stop) == cudaSuccess) {
            auto& t = ctx->prof_totals[r.name];
            t.first += ms;
            t.second += 1;
            float t0 = 0.f;
            if (tl && cudaEventElapsedTime(&t0, ctx->prof_ref, r.start) == cudaSuccess) fprintf(tl, "%s,%p,%.4f,%.4f\n", r.name, (void*)r.stream, t0, ms);
        } else {
            cudaGetLastError();
        }
        if (r.start != ctx->prof_ref) ctx->event_pool.push_back(r.start);  // the reference event stays alive
        ctx->event_pool.push_back(r.stop);
    }
    if (tl) fclose(tl);
    ctx->prof_pending.clear();
}

__global__ void mont_convert_kernel(const uint4* in, uint4* out, size_t n, int to_mont) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr v = Fr::load(in + 2 * i);
    if (to_mont) { v.reduce_once(); v = v.to_mont(); } else v = v.from_mont();
    v.store(out + 2 * i);
}

static int mont_convert(zkw_ctx* ctx, const uint64_t* in, uint64_t* out, size_t n, int to_mont) {
    if (n == 0) return ZKW_OK;
    ZKW_TRY(ensure_buffer(ctx, ctx->io_a, n * 32));
    ZKW_CUDA(ctx, cudaMemcpyAsync(ctx->io_a.ptr, in, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    { ProfScope ps_(ctx, "mont_convert_kernel"); mont_convert_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>((const uint4*)ctx->io_a.ptr, (uint4*)ctx->io_a.ptr, n, to_mont); }
    ZKW_LAUNCHED(ctx);
    ZKW_CUDA(ctx, cudaMemcpyAsync(out, ctx->io_a.ptr, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    ZKW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKW_OK;
}

static int free_buffer(DeviceBuffer& b) {
    if (b.ptr) cudaFree(b.ptr);
    b.ptr = nullptr;
    b.bytes = 0;
    return 0;
}

}  // namespace zkw

using namespace zkw;

#define CTX_ENTER(ctx)                                             \
    if (!(ctx)) return ZKW_ERR_INVALID;                            \
    ZKW_CUDA(ctx, cudaSetDevice((ctx)->device))

// host round trip shared by the host-pointer transforms
template <class F>
static int host_transform(zkw_ctx* ctx, const uint64_t* src, size_t n_in, uint64_t* dst, size_t n_out, F&& run) {
    ZKW_TRY(ensure_buffer(ctx, ctx->io_a, n_in * 32));
    ZKW_CUDA(ctx, cudaMemcpyAsync(ctx->io_a.ptr, src, n_in * 32, cudaMemcpyHostToDevice, ctx->stream));
    uint64_t* out_dev = (uint64_t*)ctx->io_a.ptr;
    if (n_out != n_in) {
        ZKW_TRY(ensure_buffer(ctx, ctx->io_b, n_out * 32));
        out_dev = (uint64_t*)ctx->io_b.ptr;
    }
    ZKW_TRY(run((uint64_t*)ctx->io_a.ptr, out_dev));
    ZKW_CUDA(ctx, cudaMemcpyAsync(dst, out_dev, n_out * 32, cudaMemcpyDeviceToHost, ctx->stream));
    ZKW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKW_OK;
}


// ---- device self-test of the hand-written field arithmetic variants -------------------------------------------------------
// sqr_lazy(a) against mul_lazy(a, a) limb for limb, over the whole lazy range [0, 2m) of both fields: pseudo-random values plus
// the corners 0, 1, m - 1, m, m + 1, 2m - 1; mul2_lazy against two separate products.
template <class F>
__device__ F selftest_value(uint64_t seed, unsigned i) {
    F a;
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL * (uint64_t)(i + 1);
    for (int j = 0; j < 4; j++) {
        z += 0x9E3779B97F4A7C15ULL;
        uint64_t w = z;
        w = (w ^ (w >> 30)) * 0xBF58476D1CE4E5B9ULL;
        w = (w ^ (w >> 27)) * 0x94D049BB133111EBULL;
        w ^= w >> 31;
        a.l[2 * j] = (uint32_t)w;
        a.l[2 * j + 1] = (uint32_t)(w >> 32);
    }
    a.l[7] &= 0x7FFFFFFFu;                 // < 2^255 < 4m
    a = a.reduced_2m();                    // [0, 2m)
    if (i < 6) {
        F m1 = F::zero() - F::one();        // m - 1
        F one = F::zero();
        one.l[0] = 1;                       // the integer 1 (not Montgomery one): any residue will do
        if (i == 0) a = F::zero();
        else if (i == 1) a = one;
        else if (i == 2) a = m1;
        else if (i == 3) a = F::add_raw(m1, one);                       // m
        else if (i == 4) a = F::add_raw(F::add_raw(m1, one), one);      // m + 1
        else a = F::add_raw(F::add_raw(m1, m1), one);                   // 2m - 1
    }
    return a;
}
template <class F>
__device__ bool selftest_sqr_one(uint64_t seed, unsigned i) {
    const F a = selftest_value<F>(seed, i);
    const F s = F::sqr_lazy(a), p = F::mul_lazy(a, a);
    bool ok = true;
    for (int j = 0; j < 8; j++) ok = ok && s.l[j] == p.l[j];
    return ok;
}
// mul2_lazy(a, b, c, d) = a b + c d (one reduction for two products) against the two separate products, and neg_2m; the first
// 6^2 indices pair the corner values with each other (2m - 1 four times is the bound's worst case), index 36 uses c = 2m (what
// neg_2m returns for zero)
template <class F>
__device__ bool selftest_mul2_one(uint64_t seed, unsigned i) {
    F a, b, c, d;
    if (i < 36) {
        a = selftest_value<F>(seed, i % 6); b = selftest_value<F>(seed, i / 6);
        c = selftest_value<F>(seed, 5 - i % 6); d = selftest_value<F>(seed, i / 6);
        if (i == 35) { a = b = c = d = selftest_value<F>(seed, 5); }
    } else {
        a = selftest_value<F>(seed, 4 * i); b = selftest_value<F>(seed, 4 * i + 1);
        c = selftest_value<F>(seed, 4 * i + 2); d = selftest_value<F>(seed, 4 * i + 3);
        if (i == 36) c = F::zero().neg_2m();
    }
    const F got = F::mul2_lazy(a, b, c, d).reduced_2m().normalized();
    const F want = F::mul_lazy(a, b).normalized() + F::mul_lazy(c.reduced_2m(), d).normalized();
    const F neg = F::add_raw(a, a.neg_2m()).reduced_2m().normalized();     // a + (2m - a) = 2m -> 0
    bool ok = true;
    for (int j = 0; j < 8; j++) ok = ok && got.l[j] == want.l[j] && neg.l[j] == 0;
    return ok;
}
__global__ void selftest_sqr_kernel(uint64_t seed, unsigned count, unsigned* mismatches) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    if (!selftest_sqr_one<zkw::Fr>(seed, i)) atomicAdd(&mismatches[0], 1u);
    if (!selftest_sqr_one<zkw::Fq>(seed ^ 0x5555, i)) atomicAdd(&mismatches[1], 1u);
    if (!selftest_mul2_one<zkw::Fr>(seed ^ 0x3333, i)) atomicAdd(&mismatches[0], 1u);
    if (!selftest_mul2_one<zkw::Fq>(seed ^ 0x7777, i)) atomicAdd(&mismatches[1], 1u);
}

extern "C" {

const char* zkw_strerror(int status) {
    switch (status) {
        case ZKW_OK: return "ok";
        case ZKW_ERR_NO_DEVICE: return "no usable CUDA device (there is no CPU fallback)";
        case ZKW_ERR_CUDA: return "CUDA runtime error";
        case ZKW_ERR_INVALID: return "invalid argument";
        case ZKW_ERR_OOM: return "out of device memory";
        case ZKW_ERR_STATE: return "invalid state (SRS not loaded?)";
        case ZKW_ERR_UNSUPPORTED: return "unsupported circuit shape";
        case ZKW_ERR_SIGNATURE: return "signature does not verify (the ECDSA circuit has no satisfying assignment)";
        default: return "unknown status";
    }
}

int zkw_ctx_create(int device, zkw_ctx** out) {
    if (!out) return ZKW_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) { cudaGetLastError(); return ZKW_ERR_NO_DEVICE; }
    if (device < 0 || device >= count) return ZKW_ERR_NO_DEVICE;
    zkw_ctx* ctx = new zkw_ctx();
    ctx->device = device;
    e = cudaSetDevice(device);
    // The main stream carries the short, latency-critical kernels between commitments (scans, divisions,
    // lookups); the MSM lanes and the transform stream carry long throughput-bound grids.  Without priorities a
    // one-CTA scan queued behind a 2000-CTA bucket accumulation waits for the whole grid.
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, zkw::stream_priority(2));
    if (e != cudaSuccess) { int rc = set_cuda_error(ctx, e, "ctx init"); delete ctx; return rc; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    // experiment knob (MSM traffic): L2 fetch granularity in bytes (32 / 64 / 128) for this device
    if (const char* g = getenv("ZKW_L2_FETCH_GRANULARITY")) {
        const int v = atoi(g);
        if (v == 32 || v == 64 || v == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)v);
    }
    const char* env = getenv("ZKW_MSM_WINDOW_BITS");
    if (env) ctx->msm_window_bits = atoi(env);
    env = getenv("ZKW_MSM_PRECOMPUTE");
    if (env) ctx->msm_precompute = atoi(env);
    env = getenv("ZKW_MSM_BINNED_SORT");
    if (env) ctx->msm_binned_sort = atoi(env);
    env = getenv("ZKW_MSM_ZERO_COPY_OUT");
    if (env) ctx->msm_zero_copy_out = atoi(env);
    env = getenv("ZKW_MSM_BINNED_MIN_ENTRIES");
    if (env) ctx->msm_binned_min_entries = atoi(env);
    *out = ctx;
    return ZKW_OK;
}

void zkw_ctx_destroy(zkw_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    // drain every stream of the context (an early error return may have left work queued on the aux stream or the
    // lanes) and read the timing events back while the streams they were recorded on still exist
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < zkw_ctx::kAuxStreams; i++) if (ctx->aux_stream[i]) cudaStreamSynchronize(ctx->aux_stream[i]);
    for (int i = 0; i < zkw_ctx::kMsmLanes; i++)
        if (ctx->lane_stream[i]) cudaStreamSynchronize(ctx->lane_stream[i]);
    profile_collect(ctx);
    for (auto& kv : ctx->twiddles) free_buffer(kv.second);
    for (auto& kv : ctx->staged_twiddles) free_buffer(kv.second);
    free_buffer(ctx->ntt_scratch); free_buffer(ctx->msm_ws);
    for (int i = 0; i < zkw_ctx::kAuxStreams; i++) {
        free_buffer(ctx->ntt_scratch_aux[i]);
        if (ctx->aux_stream[i]) cudaStreamDestroy(ctx->aux_stream[i]);
        if (ctx->aux_join[i]) cudaEventDestroy(ctx->aux_join[i]);
    }
    if (ctx->aux_fork) cudaEventDestroy(ctx->aux_fork);
    free_buffer(ctx->io_a); free_buffer(ctx->io_b); free_buffer(ctx->io_c); free_buffer(ctx->ptr_table); free_buffer(ctx->arena);
    msm_free_basis(ctx->bases[0]); msm_free_basis(ctx->bases[1]);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    for (int i = 0; i < zkw_ctx::kMsmLanes; i++) {
        free_buffer(ctx->lane_ws[i]);
        if (ctx->lane_pinned[i]) cudaFreeHost(ctx->lane_pinned[i]);
        if (i && ctx->lane_stream[i]) cudaStreamDestroy(ctx->lane_stream[i]);
        if (ctx->lane_done[i]) cudaEventDestroy(ctx->lane_done[i]);
    }
    if (ctx->fork_event) cudaEventDestroy(ctx->fork_event);
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int zkw_ctx_sync(zkw_ctx* ctx) {
    CTX_ENTER(ctx);
    ZKW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKW_OK;
}

void* zkw_ctx_stream(zkw_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
const char* zkw_last_cuda_error(zkw_ctx* ctx) { return ctx ? ctx->last_err.c_str() : ""; }
uint64_t zkw_ctx_launch_count(zkw_ctx* ctx) { return ctx ? ctx->launches : 0; }

int zkw_msm_window_bits(zkw_ctx* ctx, size_t n) {
    if (!ctx) return ZKW_ERR_INVALID;
    return zkw::msm_window_bits(ctx, n);
}

int zkw_msm_config(zkw_ctx* ctx, int window_bits, int precompute) {
    if (!ctx || window_bits < 0 || window_bits > 24 || window_bits == 1) return ZKW_ERR_INVALID;
    ctx->msm_window_bits = window_bits;
    ctx->msm_precompute = precompute ? 1 : 0;
    return ZKW_OK;
}

int zkw_profile_enable(zkw_ctx* ctx, int on) {
    if (!ctx) return ZKW_ERR_INVALID;
    profile_collect(ctx);
    ctx->profiling = on != 0;
    return ZKW_OK;
}
int zkw_profile_filter(zkw_ctx* ctx, const char* kernel_or_null) {
    if (!ctx) return ZKW_ERR_INVALID;
    profile_collect(ctx);
    ctx->prof_filter = kernel_or_null ? kernel_or_null : "";
    return ZKW_OK;
}
int zkw_profile_reset(zkw_ctx* ctx) {
    if (!ctx) return ZKW_ERR_INVALID;
    profile_collect(ctx);
    ctx->prof_totals.clear();
    return ZKW_OK;
}
int zkw_profile_read(zkw_ctx* ctx, const char* kernel, double* total_ms, uint64_t* launches) {
    if (!ctx || !kernel) return ZKW_ERR_INVALID;
    profile_collect(ctx);
    auto it = ctx->prof_totals.find(kernel);
    if (total_ms) *total_ms = it == ctx->prof_totals.end() ? 0.0 : it->second.first;
    if (launches) *launches = it == ctx->prof_totals.end() ? 0 : it->second.second;
    return ZKW_OK;
}
int zkw_profile_names(zkw_ctx* ctx, char* buf, size_t cap) {
    if (!ctx || !buf || cap == 0) return ZKW_ERR_INVALID;
    profile_collect(ctx);
    std::string all;
    for (auto& kv : ctx->prof_totals) { if (!all.empty()) all += ","; all += kv.first; }
    if (all.size() + 1 > cap) return ZKW_ERR_INVALID;
    memcpy(buf, all.c_str(), all.size() + 1);
    return ZKW_OK;
}

int zkw_dev_alloc(zkw_ctx* ctx, size_t bytes, void** out_dev) {
    CTX_ENTER(ctx);
    if (!out_dev) return ZKW_ERR_INVALID;
    ZKW_CUDA(ctx, cudaMalloc(out_dev, bytes ? bytes : 1));
    return ZKW_OK;
}
int zkw_dev_free(zkw_ctx* ctx, void* dev) {
    CTX_ENTER(ctx);
    ZKW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ZKW_CUDA(ctx, cudaFree(dev));
    return ZKW_OK;
}
int zkw_selftest_field(zkw_ctx* ctx, uint64_t seed, unsigned count, unsigned mismatches_out[2]) {
    CTX_ENTER(ctx);
    if (!mismatches_out || !count) return ZKW_ERR_INVALID;
    unsigned* d = nullptr;
    ZKW_CUDA(ctx, cudaMalloc(&d, 8));
    ZKW_CUDA(ctx, cudaMemsetAsync(d, 0, 8, ctx->stream));
    selftest_sqr_kernel<<<(count + 127) / 128, 128, 0, ctx->stream>>>(seed, count, d);
    ZKW_LAUNCHED(ctx);
    cudaError_t e = cudaMemcpyAsync(mismatches_out, d, 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return zkw::set_cuda_error(ctx, e, "zkw_selftest_field");
    return ZKW_OK;
}

int zkw_host_alloc(zkw_ctx* ctx, size_t bytes, void** out_host) {
    CTX_ENTER(ctx);
    if (!out_host) return ZKW_ERR_INVALID;
    ZKW_CUDA(ctx, cudaHostAlloc(out_host, bytes ? bytes : 1, cudaHostAllocDefault));
    return ZKW_OK;
}
int zkw_host_free(zkw_ctx* ctx, void* host) {
    CTX_ENTER(ctx);
    ZKW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ZKW_CUDA(ctx, cudaFreeHost(host));
    return ZKW_OK;
}
int zkw_memcpy_h2d(zkw_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
    CTX_ENTER(ctx);
    ZKW_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ZKW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKW_OK;
}
int zkw_memcpy_d2h(zkw_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
    CTX_ENTER(ctx);
    ZKW_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    ZKW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKW_OK;
}

// ---- SRS ----------------------------------------------------------------------------------------
static int load_basis(zkw_ctx* ctx, MsmBasis& b, const uint64_t* host, size_t n) {
    msm_free_basis(b);
    if (!host) return ZKW_OK;
    ZKW_CUDA(ctx, cudaMalloc((void**)&b.points, n * 64));
    b.n = n;
    ZKW_CUDA(ctx, cudaMemcpyAsync(b.points, host, n * 64, cudaMemcpyHostToDevice, ctx->stream));
    ZKW_TRY(msm_prepare_basis(ctx, b));
    ZKW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKW_OK;
}

int zkw_srs_load(zkw_ctx* ctx, const uint64_t* g, const uint64_t* g_lagrange, size_t n) {
    CTX_ENTER(ctx);
    if (!g || n == 0) return ZKW_ERR_INVALID;
    ZKW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ZKW_TRY(load_basis(ctx, ctx->bases[ZKW_BASES_G], g, n));
    ZKW_TRY(load_basis(ctx, ctx->bases[ZKW_BASES_G_LAGRANGE], g_lagrange, n));
    return ZKW_OK;
}

// Device-resident variant: adopt copies of device arrays (used by the prover when the SRS was
// generated on the device).
int zkw_srs_load_dev(zkw_ctx* ctx, const uint64_t* g_dev, const uint64_t* g_lagrange_dev, size_t n) {
    CTX_ENTER(ctx);
    if (!g_dev || n == 0) return ZKW_ERR_INVALID;
    ZKW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const uint64_t* srcs[2] = {g_dev, g_lagrange_dev};
    for (int i = 0; i < 2; i++) {
        MsmBasis& b = ctx->bases[i];
        msm_free_basis(b);
        if (!srcs[i]) continue;
        ZKW_CUDA(ctx, cudaMalloc((void**)&b.points, n * 64));
        b.n = n;
        ZKW_CUDA(ctx, cudaMemcpyAsync(b.points, srcs[i], n * 64, cudaMemcpyDeviceToDevice, ctx->stream));
        ZKW_TRY(msm_prepare_basis(ctx, b));
    }
    ZKW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKW_OK;
}

int zkw_fr_to_mont(zkw_ctx* ctx, const uint64_t* canonical, uint64_t* out, size_t n) {
    CTX_ENTER(ctx);
    if ((!canonical || !out) && n) return ZKW_ERR_INVALID;
    return mont_convert(ctx, canonical, out, n, 1);
}
int zkw_fr_from_mont(zkw_ctx* ctx, const uint64_t* mont, uint64_t* out, size_t n) {
    CTX_ENTER(ctx);
    if ((!mont || !out) && n) return ZKW_ERR_INVALID;
    return mont_convert(ctx, mont, out, n, 0);
}

int zkw_srs_setup(zkw_ctx* ctx, unsigned k, const uint64_t tau[4]) {
    CTX_ENTER(ctx);
    if (!tau) return ZKW_ERR_INVALID;
    ZKW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return srs_setup(ctx, k, tau);
}

int zkw_srs_get(zkw_ctx* ctx, int which_bases, uint64_t* out_xy, size_t n) {
    CTX_ENTER(ctx);
    if (which_bases != ZKW_BASES_G && which_bases != ZKW_BASES_G_LAGRANGE) return ZKW_ERR_INVALID;
    MsmBasis& b = ctx->bases[which_bases];
    if (!b.points) return ZKW_ERR_STATE;
    if (!out_xy || n > b.n) return ZKW_ERR_INVALID;
    ZKW_CUDA(ctx, cudaMemcpyAsync(out_xy, b.points, n * 64, cudaMemcpyDeviceToHost, ctx->stream));
    ZKW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKW_OK;
}

int zkw_g1_fixed_base_mul(zkw_ctx* ctx, const uint64_t* scalars, size_t n, uint64_t* out_xy) {
    CTX_ENTER(ctx);
    if ((!scalars || !out_xy) && n) return ZKW_ERR_INVALID;
    if (n == 0) return ZKW_OK;
    ZKW_TRY(ensure_buffer(ctx, ctx->io_a, n * 32));
    ZKW_TRY(ensure_buffer(ctx, ctx->io_b, n * 64));
    ZKW_CUDA(ctx, cudaMemcpyAsync(ctx->io_a.ptr, scalars, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    ZKW_TRY(fixed_base_mul_dev(ctx, (const uint64_t*)ctx->io_a.ptr, n, (uint64_t*)ctx->io_b.ptr));
    ZKW_CUDA(ctx, cudaMemcpyAsync(out_xy, ctx->io_b.ptr, n * 64, cudaMemcpyDeviceToHost, ctx->stream));
    ZKW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKW_OK;
}

// ---- MSM ------------------------------------------------------------------------------------------
int zkw_msm_bn254_g1_dev(zkw_ctx* ctx, int which_bases, const uint64_t* bases_dev, const uint64_t* scalars_dev, size_t n,
                         uint64_t* out_xyz_dev) {
    CTX_ENTER(ctx);
    if (!scalars_dev || !out_xyz_dev) return ZKW_ERR_INVALID;
    uint64_t out[12];
    ZKW_TRY(msm_run(ctx, which_bases, bases_dev, scalars_dev, n, out));
    ZKW_CUDA(ctx, cudaMemcpyAsync(out_xyz_dev, out, 96, cudaMemcpyHostToDevice, ctx->stream));
    ZKW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKW_OK;
}

// scalars on the device, result on the host: what the prover pipeline uses (commitments go into
// the transcript on the host)
int zkw_msm_bn254_g1_dev_to_host(zkw_ctx* ctx, int which_bases, const uint64_t* bases_dev, const uint64_t* scalars_dev,
                                 size_t n, uint64_t out_xyz[12]) {
    CTX_ENTER(ctx);
    if (!scalars_dev || !out_xyz) return ZKW_ERR_INVALID;
    return msm_run(ctx, which_bases, bases_dev, scalars_dev, n, out_xyz);
}

int zkw_msm_bn254_g1(zkw_ctx* ctx, int which_bases, const uint64_t* bases, const uint64_t* scalars, size_t n,
                     uint64_t out_xyz[12]) {
    CTX_ENTER(ctx);
    if (!scalars || !out_xyz) return ZKW_ERR_INVALID;
    if (which_bases == ZKW_BASES_CALLER && !bases) return ZKW_ERR_INVALID;
    ZKW_TRY(ensure_buffer(ctx, ctx->io_a, n * 32 + 32));
    ZKW_CUDA(ctx, cudaMemcpyAsync(ctx->io_a.ptr, scalars, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    const uint64_t* bases_dev = nullptr;
    if (which_bases == ZKW_BASES_CALLER) {
        ZKW_TRY(ensure_buffer(ctx, ctx->io_b, n * 64 + 64));
        ZKW_CUDA(ctx, cudaMemcpyAsync(ctx->io_b.ptr, bases, n * 64, cudaMemcpyHostToDevice, ctx->stream));
        bases_dev = (const uint64_t*)ctx->io_b.ptr;
    }
    return msm_run(ctx, which_bases, bases_dev, (const uint64_t*)ctx->io_a.ptr, n, out_xyz);
}

int zkw_g1_batch_normalize(zkw_ctx* ctx, const uint64_t* xyz, size_t m, uint64_t* out_xy) {
    CTX_ENTER(ctx);
    if ((!xyz || !out_xy) && m) return ZKW_ERR_INVALID;
    if (m == 0) return ZKW_OK;
    ZKW_TRY(ensure_buffer(ctx, ctx->io_a, m * 96));
    ZKW_TRY(ensure_buffer(ctx, ctx->io_b, m * 64));
    ZKW_CUDA(ctx, cudaMemcpyAsync(ctx->io_a.ptr, xyz, m * 96, cudaMemcpyHostToDevice, ctx->stream));
    ZKW_TRY(g1_batch_normalize_dev(ctx, (const uint64_t*)ctx->io_a.ptr, m, (uint64_t*)ctx->io_b.ptr));
    ZKW_CUDA(ctx, cudaMemcpyAsync(out_xy, ctx->io_b.ptr, m * 64, cudaMemcpyDeviceToHost, ctx->stream));
    ZKW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKW_OK;
}

// ---- NTT --------------------------------------------------------------------------------------------
int zkw_ntt_bn254_fr_dev(zkw_ctx* ctx, uint64_t* a_dev, unsigned log_n, const uint64_t omega[4], const uint64_t* scale_or_null) {
    CTX_ENTER(ctx);
    if (!a_dev || !omega) return ZKW_ERR_INVALID;
    uint64_t s3[12];
    if (scale_or_null) for (int i = 0; i < 3; i++) memcpy(s3 + 4 * i, scale_or_null, 32);
    return ntt_run(ctx, a_dev, log_n, a_dev, log_n, omega, false, scale_or_null ? s3 : nullptr);
}

int zkw_ntt_bn254_fr(zkw_ctx* ctx, uint64_t* a, unsigned log_n, const uint64_t omega[4], const uint64_t* scale_or_null) {
    CTX_ENTER(ctx);
    if (!a || !omega || log_n > 28) return ZKW_ERR_INVALID;
    const size_t n = (size_t)1 << log_n;
    return host_transform(ctx, a, n, a, n, [&](uint64_t* in, uint64_t*) { return zkw_ntt_bn254_fr_dev(ctx, in, log_n, omega, scale_or_null); });
}

int zkw_lagrange_to_coeff_dev(zkw_ctx* ctx, uint64_t* a_dev, unsigned k) {
    CTX_ENTER(ctx);
    if (!a_dev || k > 28) return ZKW_ERR_INVALID;
    DomainConsts dc;
    domain_consts(k, k, &dc);
    return ntt_run(ctx, a_dev, k, a_dev, k, dc.omega_inv, false, dc.n_scale3);
}
int zkw_lagrange_to_coeff(zkw_ctx* ctx, uint64_t* a, unsigned k) {
    CTX_ENTER(ctx);
    if (!a || k > 28) return ZKW_ERR_INVALID;
    const size_t n = (size_t)1 << k;
    return host_transform(ctx, a, n, a, n, [&](uint64_t* in, uint64_t*) { return zkw_lagrange_to_coeff_dev(ctx, in, k); });
}
int zkw_coeff_to_lagrange_dev(zkw_ctx* ctx, uint64_t* a_dev, unsigned k) {
    CTX_ENTER(ctx);
    if (!a_dev || k > 28) return ZKW_ERR_INVALID;
    DomainConsts dc;
    domain_consts(k, k, &dc);
    return ntt_run(ctx, a_dev, k, a_dev, k, dc.omega, false, nullptr);
}
int zkw_coeff_to_lagrange(zkw_ctx* ctx, uint64_t* a, unsigned k) {
    CTX_ENTER(ctx);
    if (!a || k > 28) return ZKW_ERR_INVALID;
    const size_t n = (size_t)1 << k;
    return host_transform(ctx, a, n, a, n, [&](uint64_t* in, uint64_t*) { return zkw_coeff_to_lagrange_dev(ctx, in, k); });
}
int zkw_coeff_to_extended_dev(zkw_ctx* ctx, const uint64_t* coeffs_dev, unsigned k, unsigned ext_k, uint64_t* out_dev) {
    CTX_ENTER(ctx);
    if (!coeffs_dev || !out_dev || ext_k > 28 || k > ext_k || (const void*)coeffs_dev == (const void*)out_dev) return ZKW_ERR_INVALID;
    DomainConsts dc;
    domain_consts(k, ext_k, &dc);
    return ntt_run(ctx, coeffs_dev, k, out_dev, ext_k, dc.ext_omega, true, nullptr);
}
int zkw_coeff_to_extended(zkw_ctx* ctx, const uint64_t* coeffs, unsigned k, unsigned ext_k, uint64_t* out) {
    CTX_ENTER(ctx);
    if (!coeffs || !out || ext_k > 28 || k > ext_k) return ZKW_ERR_INVALID;
    const size_t n_in = (size_t)1 << k, n_out = (size_t)1 << ext_k;
    ZKW_TRY(ensure_buffer(ctx, ctx->io_a, n_in * 32));
    ZKW_TRY(ensure_buffer(ctx, ctx->io_b, n_out * 32));
    ZKW_CUDA(ctx, cudaMemcpyAsync(ctx->io_a.ptr, coeffs, n_in * 32, cudaMemcpyHostToDevice, ctx->stream));
    ZKW_TRY(zkw_coeff_to_extended_dev(ctx, (const uint64_t*)ctx->io_a.ptr, k, ext_k, (uint64_t*)ctx->io_b.ptr));
    ZKW_CUDA(ctx, cudaMemcpyAsync(out, ctx->io_b.ptr, n_out * 32, cudaMemcpyDeviceToHost, ctx->stream));
    ZKW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKW_OK;
}
int zkw_extended_to_coeff_dev(zkw_ctx* ctx, uint64_t* a_dev, unsigned ext_k) {
    CTX_ENTER(ctx);
    if (!a_dev || ext_k > 28) return ZKW_ERR_INVALID;
    DomainConsts dc;
    domain_consts(ext_k, ext_k, &dc);
    return ntt_run(ctx, a_dev, ext_k, a_dev, ext_k, dc.ext_omega_inv, false, dc.ext_scale3);
}
int zkw_extended_to_coeff(zkw_ctx* ctx, uint64_t* a, unsigned ext_k) {
    CTX_ENTER(ctx);
    if (!a || ext_k > 28) return ZKW_ERR_INVALID;
    const size_t n = (size_t)1 << ext_k;
    return host_transform(ctx, a, n, a, n, [&](uint64_t* in, uint64_t*) { return zkw_extended_to_coeff_dev(ctx, in, ext_k); });
}

// ---- quotient ------------------------------------------------------------------------------------------
int zkw_quotient_ecdsa_dev(zkw_ctx* ctx, const zkw_quotient_inputs* in, uint64_t* h_ext_dev) {
    CTX_ENTER(ctx);
    if (!in || !h_ext_dev) return ZKW_ERR_INVALID;
    return quotient_run(ctx, in, h_ext_dev);
}

int zkw_quotient_ecdsa(zkw_ctx* ctx, const zkw_quotient_inputs* in, uint64_t* h_ext) {
    CTX_ENTER(ctx);
    if (!in || !h_ext) return ZKW_ERR_INVALID;
    const zkw_circuit_shape& sh = in->shape;
    if (sh.ext_k > 28) return ZKW_ERR_INVALID;
    const unsigned A = sh.num_advice, L = sh.num_lookup_advice, F = sh.num_fixed;
    const unsigned ncols = F + A + L, nsets = zkw_shape_perm_sets(&sh), nlk = zkw_shape_lookups(&sh);
    if (L == 0 && !in->q_lookup) return ZKW_ERR_UNSUPPORTED;
    const size_t en = (size_t)1 << sh.ext_k, vb = en * 32;
    // stage every input coset in one device arena
    std::vector<const uint64_t*> hosts;
    auto add = [&](const uint64_t* const* t, unsigned cnt) { for (unsigned i = 0; i < cnt; i++) hosts.push_back(t ? t[i] : nullptr); };
    add(in->advice, A + L); add(in->constants, F); add(in->q_enable, A); add(in->sigma, ncols);
    add(in->perm_z, nsets); add(in->lookup_z, nlk); add(in->lookup_a, nlk); add(in->lookup_s, nlk);
    hosts.push_back(in->table); hosts.push_back(in->l0); hosts.push_back(in->l_last); hosts.push_back(in->l_active);
    if (L == 0) hosts.push_back(in->q_lookup);
    for (auto p : hosts) if (!p) return ZKW_ERR_INVALID;
    ZKW_TRY(ensure_buffer(ctx, ctx->io_c, (hosts.size() + 1) * vb));
    char* arena = (char*)ctx->io_c.ptr;
    std::vector<const uint64_t*> devs(hosts.size());
    for (size_t i = 0; i < hosts.size(); i++) {
        devs[i] = (const uint64_t*)(arena + i * vb);
        ZKW_CUDA(ctx, cudaMemcpyAsync((void*)devs[i], hosts[i], vb, cudaMemcpyHostToDevice, ctx->stream));
    }
    uint64_t* h_dev = (uint64_t*)(arena + hosts.size() * vb);
    zkw_quotient_inputs d = *in;
    size_t o = 0;
    d.advice = devs.data() + o; o += A + L;
    d.constants = devs.data() + o; o += F;
    d.q_enable = devs.data() + o; o += A;
    d.sigma = devs.data() + o; o += ncols;
    d.perm_z = devs.data() + o; o += nsets;
    d.lookup_z = devs.data() + o; o += nlk;
    d.lookup_a = devs.data() + o; o += nlk;
    d.lookup_s = devs.data() + o; o += nlk;
    d.table = devs[o++]; d.l0 = devs[o++]; d.l_last = devs[o++]; d.l_active = devs[o++];
    d.q_lookup = (L == 0) ? devs[o++] : nullptr;
    ZKW_TRY(quotient_run(ctx, &d, h_dev));
    ZKW_CUDA(ctx, cudaMemcpyAsync(h_ext, h_dev, vb, cudaMemcpyDeviceToHost, ctx->stream));
    ZKW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKW_OK;
}

}  // extern "C"
