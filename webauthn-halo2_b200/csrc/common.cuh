// common.cuh — context object and launch plumbing shared by the kernels behind include/zkw_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <map>
#include <string>
#include <vector>
#include <array>
#include "../../include/zkw_b200.h"
#include "curve.cuh"
#include "inverse.cuh"

namespace zkw {

struct TwiddleKey {
    std::array<uint64_t, 4> omega;
    unsigned log_n;
    bool operator<(const TwiddleKey& o) const {
        if (log_n != o.log_n) return log_n < o.log_n;
        return omega < o.omega;
    }
};

struct StagedTwiddleKey {
    std::array<uint64_t, 4> omega;
    unsigned log_n, s0, B;
    bool operator<(const StagedTwiddleKey& o) const {
        if (log_n != o.log_n) return log_n < o.log_n;
        if (s0 != o.s0) return s0 < o.s0;
        if (B != o.B) return B < o.B;
        return omega < o.omega;
    }
};

struct DeviceBuffer {
    void* ptr = nullptr;
    size_t bytes = 0;
};

// Precomputed window multiples of a fixed basis: table[w*n + i] = 2^(c*w) * P_i (affine).
struct MsmBasis {
    uint64_t* points = nullptr;  // n affine points as loaded
    uint64_t* table = nullptr;   // windows*n affine points, or nullptr when precomputation is off
    size_t n = 0;
    int c = 0, windows = 0;
};

}  // namespace zkw

struct zkw_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string last_err;
    uint64_t launches = 0;
    int sm_count = 148;

    // SRS residency
    zkw::MsmBasis bases[2];
    // MSM tuning: window bits (0 = automatic) and whether fixed bases get window tables
    int msm_window_bits = 0;
    int msm_precompute = 1;
    // bucket sort of the MSM entries: binned (shared-memory counting, msm.cu) unless ZKW_MSM_BINNED_SORT=0; MSMs with
    // fewer entries than msm_binned_min_entries keep the direct three-kernel sort (fewer launches)
    int msm_binned_sort = 1;
    int msm_binned_min_entries = 1 << 16;
    int msm_zero_copy_out = 1;           // MSM results written by the kernel into page-locked host memory (ZKW_MSM_ZERO_COPY_OUT=0: D2H copy)
    bool ntt_attr_set = false;           // NTT pass kernel's dynamic shared memory opt-in done
    bool msm_attr_set = false;           // accumulate kernel's dynamic shared memory opt-in done

    // twiddle tables: omega^i for i < 2^(log_n-1), keyed by (omega, log_n)
    std::map<zkw::TwiddleKey, zkw::DeviceBuffer> twiddles;
    // per-(omega, log_n, pass) compact copies of a pass's last-stage twiddles, laid out for one bulk copy per tile
    std::map<zkw::StagedTwiddleKey, zkw::DeviceBuffer> staged_twiddles;
    // reusable scratch areas (grown on demand, never shrunk)
    // prover: transforms that overlap the main stream's MSMs run on an auxiliary stream (up to three, round robin, with
    // ZKW_AUX_STREAMS: measured, no gain — prover.cu create_proof_impl)
    static constexpr int kAuxStreams = 3;
    zkw::DeviceBuffer ntt_scratch, ntt_scratch_aux[kAuxStreams];
    cudaStream_t aux_stream[kAuxStreams] = {nullptr};
    cudaEvent_t aux_fork = nullptr, aux_join[kAuxStreams] = {nullptr};
    int aux_count = 0, aux_next = 0;     // streams in use (ZKW_AUX_STREAMS, default 1) and the round-robin cursor
    zkw::DeviceBuffer msm_ws;
    // MSM lanes: lane 0 runs on `stream`; lanes 1.. own a side stream so that the latency-bound tails of
    // one MSM overlap the accumulation of the next (msm_run_batch)
    static constexpr int kMsmLanes = 6;
    cudaStream_t lane_stream[kMsmLanes] = {nullptr};
    cudaEvent_t lane_done[kMsmLanes] = {nullptr};
    cudaEvent_t fork_event = nullptr;
    zkw::DeviceBuffer lane_ws[kMsmLanes];
    void* lane_pinned[kMsmLanes] = {nullptr};
    size_t lane_pinned_bytes[kMsmLanes] = {0};
    int lane_groups[kMsmLanes] = {0}, lane_c[kMsmLanes] = {0};
    zkw::DeviceBuffer io_a, io_b, io_c;  // staging for the host-pointer entry points
    zkw::DeviceBuffer ptr_table;         // device copy of the quotient pointer tables
    zkw::DeviceBuffer arena;             // per-proof scratch arena (prover.cu), grown to the high-water mark
    size_t arena_off = 0, arena_virtual = 0, arena_need = 0;
    void* pinned = nullptr;              // small pinned host area for results
    size_t pinned_bytes = 0;

    // optional per-kernel timing (CUDA events on the ctx stream), off by default
    bool profiling = false;
    std::string prof_filter;          // non-empty: only launches of this kernel are timed (zkw_profile_filter)
    struct ProfRec { const char* name; cudaEvent_t start, stop; cudaStream_t stream; };
    cudaEvent_t prof_ref = nullptr;      // timeline origin (ZKW_TIMELINE)
    std::vector<ProfRec> prof_pending;
    std::map<std::string, std::pair<double, uint64_t>> prof_totals;  // name -> (ms, launches)
    std::vector<cudaEvent_t> event_pool;
};

namespace zkw {

int set_cuda_error(zkw_ctx* ctx, cudaError_t e, const char* what);
int stream_priority(int level);  // 0 = background (MSM lanes), 1 = transforms, 2 = the context's main stream
int ensure_buffer(zkw_ctx* ctx, DeviceBuffer& b, size_t bytes);

#define ZKW_CUDA(ctx, call)                                            \
    do {                                                               \
        cudaError_t _e = (call);                                       \
        if (_e != cudaSuccess) return zkw::set_cuda_error(ctx, _e, #call); \
    } while (0)

#define ZKW_TRY(expr)            \
    do {                         \
        int _rc = (expr);        \
        if (_rc != ZKW_OK) return _rc; \
    } while (0)

// after every kernel launch: count it and surface launch-configuration errors
#define ZKW_LAUNCHED(ctx)                                                   \
    do {                                                                    \
        (ctx)->launches++;                                                  \
        cudaError_t _e = cudaGetLastError();                                \
        if (_e != cudaSuccess) return zkw::set_cuda_error(ctx, _e, "kernel launch"); \
    } while (0)

// RAII bracket around one kernel launch: records start/stop events when ctx->profiling is on.
struct ProfScope {
    zkw_ctx* ctx;
    cudaEvent_t stop = nullptr;
    cudaStream_t stream;
    ProfScope(zkw_ctx* c, const char* name, cudaStream_t s = nullptr);
    ~ProfScope();
};

// ---- entry points implemented per translation unit (device pointers, async on ctx->stream) ----
// ntt.cu
int ntt_get_twiddles(zkw_ctx* ctx, const uint64_t omega[4], unsigned log_n, const uint64_t** out_dev);
// generic transform: dst (2^log_n) <- NTT_omega(src'), src' = src zero-extended from 2^src_log_n with
// optional zeta^(i mod 3) pre-scaling (coset) ; optional per-(i mod 3) output scaling. src may equal dst.
int ntt_run(zkw_ctx* ctx, const uint64_t* src_dev, unsigned src_log_n, uint64_t* dst_dev, unsigned log_n,
            const uint64_t omega[4], bool coset_in, const uint64_t* scale3 /* 3*4 u64 host, or NULL */,
            cudaStream_t stream = nullptr /* default: the ctx stream */);
// msm.cu
int msm_run(zkw_ctx* ctx, int which_bases, const uint64_t* bases_dev, const uint64_t* scalars_dev, size_t n,
            uint64_t out_xyz_host[12]);
struct MsmJob { int which_bases; const uint64_t* bases_dev; const uint64_t* scalars_dev; size_t n; };
// up to zkw_ctx::kMsmLanes independent MSMs in flight at once; outs[i] = 12 u64 (x, y, 1) or Z = 0
int msm_run_batch(zkw_ctx* ctx, const MsmJob* jobs, int count, uint64_t (*outs)[12]);
// pipelined form: submit one MSM to a side lane (1 <= lane < kMsmLanes) as soon as its scalars are queued
// on the main stream, keep working on the main stream, and collect later (synchronises)
int msm_lane_submit(zkw_ctx* ctx, int lane, const MsmJob& job);
int msm_lanes_collect(zkw_ctx* ctx, const int* lanes, int count, uint64_t (*outs)[12]);
// wait for ONE lane's MSM (host blocks on that lane's event only) and fetch its result
int msm_lane_wait(zkw_ctx* ctx, int lane, uint64_t out_xyz[12]);
int msm_prepare_basis(zkw_ctx* ctx, MsmBasis& b);
int msm_window_bits(const zkw_ctx* ctx, size_t n);   // the window width an n-point MSM uses
void msm_free_basis(MsmBasis& b);
int g1_batch_normalize_dev(zkw_ctx* ctx, const uint64_t* xyz_dev, size_t m, uint64_t* out_xy_dev);
// srs.cu
int srs_setup(zkw_ctx* ctx, unsigned k, const uint64_t tau_m[4]);
int fixed_base_mul_dev(zkw_ctx* ctx, const uint64_t* scalars_dev, size_t n, uint64_t* out_xy_dev);
// quotient.cu
int quotient_run(zkw_ctx* ctx, const zkw_quotient_inputs* in /* device vectors */, uint64_t* h_ext_dev);

#ifdef __CUDACC__
// ---- TMA (bulk asynchronous copy) plumbing, used by ntt.cu (twiddle staging): one
// elected thread arms an mbarrier with the byte count and issues cp.async.bulk global -> shared; the consumers wait
// on the barrier's phase ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
#endif

// domain constants (host side, Montgomery form), see domain.cpp
struct DomainConsts {
    unsigned k, ext_k;
    uint64_t omega[4], omega_inv[4], ext_omega[4], ext_omega_inv[4];
    uint64_t zeta[4], zeta_inv[4];            // g_coset, g_coset_inv (= zeta^2)
    uint64_t n_inv[4], ext_n_inv[4];
    uint64_t ext_scale3[12];                  // 2^-ext_k * zeta^-(i mod 3), i = 0,1,2
    uint64_t n_scale3[12];                    // n^-1 replicated
    uint64_t t_evals[16][4];                  // 1/((zeta*w_ext^i)^n - 1), i < 2^(ext_k-k)
};
void domain_consts(unsigned k, unsigned ext_k, DomainConsts* out);

}  // namespace zkw
