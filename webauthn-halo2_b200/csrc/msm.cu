// msm.cu — multi-scalar multiplication over BN254 G1 for sm_100a: the device replacement of
// halo2_proofs::arithmetic::best_multiexp (KZG commit / commit_lagrange), reached from the reference
// through create_proof and keygen (halo2-circuits/src/ecc/ecdsa_p256.rs:259-260, 366-373, 416-423,
// 555-562).  Result = sum_i s_i * P_i; only its affine normalisation is canonical, and that is what
// the C ABI hands back (as Jacobian (x, y, 1)).
//
// Pippenger, re-laid-out for the GPU:
//   recode     scalars leave Montgomery form (upstream's to_repr()) and are cut into W signed
//              c-bit digits d in [-2^(c-1), 2^(c-1)]; |d| names a bucket, the sign negates the point.
//   fixed SRS  for the resident bases (g, g_lagrange) the context holds 2^(c*w) * P_i for every
//              window w, so all windows share ONE set of 2^(c-1) buckets and there is no per-window
//              doubling chain at the end ("groups" G = 1).  Caller-supplied bases use one bucket set
//              per window (G = W) and the windows are combined on the host.
//   sort       counting sort of the (window, point) entries by bucket.  Default: binned - coarse bins of 128 buckets, every
//              count and rank taken in shared memory, chunks of a bin staged by one TMA bulk copy (bin_count, bin_scan,
//              bin_scatter, fine_count, scan, fine_scatter; see "binned counting sort" below).  Small, very large or
//              oddly shaped MSMs: direct - histogram (atomics in L2), single-CTA scan, scatter.
//   accumulate the sorted entry list is cut into EQUAL runs of L = ceil(E / T) entries, T = ZKW_MSM_WAVES resident waves of
//              threads (CTAs per SM x SMs x 128 each), so every thread does the same number of mixed XYZZ additions (8M + 2S
//              each) whatever the scalar distribution (witness columns are far from uniform: zeros, bits, small limbs).
//              A run crosses bucket boundaries: thread t emits its sum for bucket b to partial slot t + b
//              (strictly increasing along the entry list, so a bucket's partials are contiguous and their
//              range follows from the bucket's offsets alone).  The next affine point is prefetched (two
//              128-bit loads per coordinate) while the current one is added.
//   combine    one thread per bucket folds its (typically 2-4) partials; buckets cut into more than kLight runs (skewed
//              scalars) are queued: up to kMedium partials one warp each, above that one CTA, the heaviest split over
//              several CTAs whose sums the last one to finish folds.
//   reduce     sum_b b * B_b by rows and columns of the bucket index (b - 1 = hi * 2^lb + lo): tree sums of
//              every row and column, then two short bit-sliced weighted sums; the final ~2c-step Horner
//              runs on the host in microseconds instead of as a latency-bound chain on the device.
//
// The kernels are integer-ALU bound: 96 algorithmic bytes per point against ~10 field products per
// window per point.  DESIGN.md carries the roofline arithmetic.
#include <algorithm>
#include "common.cuh"

namespace zkw {

// Resident accumulate CTAs per SM: 120 registers x 128 threads allow four.  ZKW_MSM_SMEM_RESERVE > 0 makes the
// kernel ask for that much dynamic shared memory it never touches, which caps the residency of accumulate CTAs
// from all MSM lanes together and keeps a CTA slot per SM free for the short kernels of other streams.  Measured
// (tools/msm_ab.py, k = 19 proof): 4 CTAs / no reserve 28.2 ms, 3 CTAs + 64 KB 28.7 ms, 2 CTAs + 96 KB 29.6 ms -
// four warps per scheduler hide the multiplier latency better than three, and that outweighs the free slot.
#ifndef ZKW_MSM_CTAS_PER_SM
#define ZKW_MSM_CTAS_PER_SM 4
#endif
#ifndef ZKW_MSM_SMEM_RESERVE
#define ZKW_MSM_SMEM_RESERVE 0
#endif
constexpr int kAccSmemReserve = ZKW_MSM_SMEM_RESERVE;
// Accumulate grid in resident waves.  One wave gives the longest runs and the fewest partials, but its CTAs live for
// the whole kernel: any slot another stream's kernel holds when the grid launches pushes accumulate CTAs into a
// second, nearly empty wave.  With W waves the runs are W times shorter and that tail is bounded by 1/W of the kernel.
// Measured (tools/msm_ab.py, same call): W = 1 / 2 / 3 / 4 -> uniform MSM 1.955 / 1.93 / 1.92 / 1.93 ms, k = 19 proof
// 27.38 / 27.47 / 27.46 / 27.75 ms: the proof is multiplier-bound either way, one wave keeps the partials fewest.
// Re-measured with the final round-2 build (binned sort, c = 17; same call, twice): W = 1 / 2 / 3 -> MSM 1.602 / 1.585 / 1.608 ms,
// proof 24.24 / 24.12 / 24.22 ms: two waves it is (the sort and combine kernels of the other lanes now find a slot mid-kernel).
#ifndef ZKW_MSM_WAVES
#define ZKW_MSM_WAVES 2
#endif
#ifndef ZKW_MSM_ACC_THREADS
#define ZKW_MSM_ACC_THREADS 128
#endif
constexpr int kAccThreads = ZKW_MSM_ACC_THREADS;
#ifndef ZKW_MSM_MIN_RUN
#define ZKW_MSM_MIN_RUN 16
#endif
constexpr int kMinRun = ZKW_MSM_MIN_RUN;        // shortest run worth a thread (small MSMs use fewer threads instead)
constexpr int kLight = ZKW_MSM_WAVES > 1 ? 16 : 8;          // partials per bucket folded by one thread; more -> queued
constexpr int kMedium = 128;        // up to this many partials a queued bucket gets one warp, above it a CTA (or several)
constexpr int kReduceThreads = 64;  // CTA size of the row / column bucket reduction (one warp per row or column)

struct MsmPlan {
    int c;            // window bits
    int windows;      // W
    int groups;       // G: 1 with window tables, W otherwise
    uint32_t nb;      // buckets per group = 2^(c-1)
    size_t n;
    size_t max_entries() const { return (size_t)windows * n; }
    size_t total_buckets() const { return (size_t)groups * nb; }
};

// device-side run plan, written by the scan kernel once the number of entries is known
struct RunPlan { uint32_t entries, run, threads, heavy, medium; };  // heavy / medium: buckets queued for msm_combine_heavy_kernel / msm_combine_medium_kernel

// ---- recode + histogram -----------------------------------------------------------------------
// the signed c-bit digits of a canonical scalar, lowest window first: bit 31 = sign, low bits = |d| in [0, 2^(c-1)];
// the scalar is shifted down by c bits per digit (funnel shifts, no indexed limb access); c <= 31
struct DigitStream {
    uint32_t l[8];
    uint32_t carry;
    __device__ __forceinline__ explicit DigitStream(const Fr& s) : carry(0) {
#pragma unroll
        for (int j = 0; j < 8; j++) l[j] = s.l[j];
    }
    __device__ __forceinline__ uint32_t next(int c) {
        uint32_t v = (l[0] & ((1u << c) - 1u)) + carry;
#pragma unroll
        for (int j = 0; j < 7; j++) l[j] = __funnelshift_r(l[j], l[j + 1], c);
        l[7] >>= c;
        if (v > (1u << (c - 1))) {  // negative digit: v - 2^c
            carry = 1;
            return 0x80000000u | ((1u << c) - v);
        }
        carry = 0;
        return v;
    }
};

__global__ void msm_recode_kernel(const uint4* __restrict__ scalars, uint32_t* __restrict__ digits,
                                  uint32_t* __restrict__ counts, size_t n, int c, int windows, int groups, uint32_t nb) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    DigitStream ds(Fr::load_nc(scalars + 2 * i).from_mont());
    for (int w = 0; w < windows; w++) {
        const uint32_t out = ds.next(c);
        digits[(size_t)w * n + i] = out;
        const uint32_t mag = out & 0x7fffffffu;
        if (mag) atomicAdd(&counts[(groups > 1 ? (size_t)w * nb : 0) + (mag - 1)], 1u);
    }
}

// ---- single-CTA exclusive scan of the bucket counts -> bucket offsets; also fixes the run length ------
// Tiles of 16384 counts (sixteen consecutive counts per thread); per tile a warp-shuffle scan of the thread
// totals, a scan of the 32 warp totals, and a running carry.
__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

__global__ void __launch_bounds__(1024) msm_scan_kernel(const uint32_t* __restrict__ counts, uint32_t* __restrict__ offsets,
                                                        RunPlan* __restrict__ plan, size_t total, uint32_t max_threads) {
    constexpr int kPer = 16;   // counts per thread and tile: four 128-bit loads in flight, one block scan per 16384 counts
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) carry = 0;
    __syncthreads();
    for (size_t base = 0; base < total; base += 1024 * kPer) {
        const size_t idx = base + kPer * (size_t)t;
        uint32_t cnt[kPer];
        if (idx + kPer <= total) {
#pragma unroll
            for (int q = 0; q < kPer / 4; q++) {
                const uint4 v = *reinterpret_cast<const uint4*>(counts + idx + 4 * q);
                cnt[4 * q] = v.x; cnt[4 * q + 1] = v.y; cnt[4 * q + 2] = v.z; cnt[4 * q + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < kPer; j++) cnt[j] = idx + j < total ? counts[idx + j] : 0u;
        }
        uint32_t tot = 0;
#pragma unroll
        for (int j = 0; j < kPer; j++) tot += cnt[j];
        const uint32_t inc = warp_inclusive_scan(tot, lane);
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) wsum[lane] = warp_inclusive_scan(wsum[lane], lane);
        __syncthreads();
        uint32_t run = carry + (warp ? wsum[warp - 1] : 0u) + inc - tot;
        if (idx + kPer <= total) {
#pragma unroll
            for (int q = 0; q < kPer / 4; q++) {
                uint4 o;
                o.x = run; run += cnt[4 * q];
                o.y = run; run += cnt[4 * q + 1];
                o.z = run; run += cnt[4 * q + 2];
                o.w = run; run += cnt[4 * q + 3];
                *reinterpret_cast<uint4*>(offsets + idx + 4 * q) = o;
            }
        } else {
#pragma unroll
            for (int j = 0; j < kPer; j++) {
                if (idx + j < total) { offsets[idx + j] = run; run += cnt[j]; }
            }
        }
        __syncthreads();
        if (t == 0) carry += wsum[31];
        __syncthreads();
    }
    if (t == 0) {
        const uint32_t e = carry;
        offsets[total] = e;
        uint32_t run = (e + max_threads - 1) / max_threads;
        if (run < (uint32_t)kMinRun) run = kMinRun;
        plan->entries = e;
        plan->run = run;
        plan->threads = (e + run - 1) / run;
        plan->heavy = 0;
        plan->medium = 0;
    }
}

// ---- scatter entries into bucket order ----------------------------------------------------------
__global__ void msm_scatter_kernel(const uint32_t* __restrict__ digits, const uint32_t* __restrict__ offsets,
                                   uint32_t* __restrict__ cursor, uint32_t* __restrict__ sorted, size_t n, int windows,
                                   int groups, uint32_t nb, int table) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)windows * n) return;
    const uint32_t d = digits[e];
    const uint32_t mag = d & 0x7fffffffu;
    if (!mag) return;
    const size_t w = e / n, i = e - w * n;
    const size_t b = (groups > 1 ? w * nb : 0) + (mag - 1);
    const uint32_t pos = offsets[b] + atomicAdd(&cursor[b], 1u);
    const uint32_t pidx = table ? (uint32_t)e : (uint32_t)i;  // table[w*n + i] = 2^(c w) P_i
    sorted[pos] = (d & 0x80000000u) | pidx;
}

// ---- binned counting sort: the same bucket order without one L2 atomic per entry ---------------------------
// The direct sort above pays two global atomics per entry (histogram, cursor): 2 x 8.4 M at 2^19 points, 230 us, and
// far worse when a witness column piles its entries on a few buckets.  Here the bucket id is split into a coarse bin
// (id >> 7) and a fine index (id & 127), and every count / rank is taken in SHARED memory:
//   bin_count    each CTA recodes a stripe of scalars and histograms the coarse bins in shared memory; one global add
//                per (CTA, bin).
//   bin_scan     one CTA: bin offsets (each bin padded to a multiple of four entries so that a chunk is a legal bulk
//                copy) and the number of 4096-entry chunks per bin.
//   bin_scatter  recodes again, ranks its entries per bin in shared memory, reserves a range per (CTA, bin) with one
//                global add and writes the entries, packed as fine | sign | point id, bin by bin.
//   fine_count   one CTA per chunk: the chunk (16 KB, contiguous) is staged into shared memory by ONE bulk copy (TMA,
//                cp.async.bulk completing on an mbarrier), histogrammed over the bin's 128 buckets in shared memory, and
//                the chunk's count per bucket is added to the global bucket count - the value the add returns is the
//                chunk's base inside the bucket, kept for the last step.
//   (msm_scan_kernel: bucket offsets and the run plan, as before)
//   fine_scatter the chunk is staged again, every entry takes its rank from a shared-memory cursor that starts at
//                offsets[bucket] + the chunk's base, and lands in its final slot.
// Skewed columns (constant stretches of a grand product, zero-heavy limbs) pile onto few shared-memory words, which the
// hardware serialises at one lane per clock: they cost less than uniform ones.
constexpr int kFineBits = 7;
constexpr int kFine = 1 << kFineBits;
constexpr int kMaxBins = 4096;
constexpr int kBinThreads = 1024;
constexpr int kMaxWindows = 32;        // msm_prepare_basis / msm_enqueue keep the binned path to at most 32 windows
constexpr int kChunk = 4096;            // entries per chunk: 16 KB of shared memory
constexpr int kChunkThreads = 512;
constexpr size_t kBinnedMaxEntries = (size_t)1 << 24;   // packed entry: 7 bits fine, 1 bit sign, 24 bits point id
// (A 64-bit entry for larger MSMs was built and measured: 2^21 points 5.58 ms either way, 2^22 points 10.50 ms binned against
// 10.23 ms with the direct sort - with 2^19 buckets the direct sort's atomics are spread thin enough.  Not kept.)

// one more entry for `key` in the shared-memory table; returns the entry's rank.  (Aggregating the lanes of a warp that
// agree on the key with match_all before the add was measured and dropped: the hardware already serialises same-address
// shared-memory atomics at one lane per clock, faster than the two per lane of spread addresses - a permuted lookup
// column's bin_scatter takes 43 us without the aggregation, 47 us with it.)
__device__ __forceinline__ uint32_t smem_rank(uint32_t* table, uint32_t key, bool active) {
    return active ? atomicAdd(&table[key], 1u) : 0u;
}

__global__ void __launch_bounds__(kBinThreads) msm_bin_count_kernel(const uint4* __restrict__ scalars, uint32_t* __restrict__ bin_counts,
                                                                    size_t n, int c, int windows, int groups, uint32_t nb, uint32_t nbins,
                                                                    uint32_t stripe) {
    __shared__ uint32_t s_cnt[kMaxBins];
    for (uint32_t b = threadIdx.x; b < nbins; b += kBinThreads) s_cnt[b] = 0;
    __syncthreads();
    for (size_t base = (size_t)blockIdx.x * stripe; base < n; base += (size_t)gridDim.x * stripe) {
        const size_t i = base + threadIdx.x;
        Fr s = Fr::zero();
        if (threadIdx.x < stripe && i < n) s = Fr::load_nc(scalars + 2 * i).from_mont();
        DigitStream ds(s);
        for (int w = 0; w < windows; w++) {
            const uint32_t mag = ds.next(c) & 0x7fffffffu;
            const uint32_t id = (groups > 1 ? (uint32_t)w * nb : 0u) + (mag - 1u);
            smem_rank(s_cnt, id >> kFineBits, mag != 0);
        }
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nbins; b += kBinThreads) {
        const uint32_t v = s_cnt[b];
        if (v) atomicAdd(&bin_counts[b], v);
    }
}

// one CTA, nbins <= 4096: bin_offsets (padded to multiples of four entries) and chunk_first, both exclusive scans with
// the total in slot [nbins]
__global__ void __launch_bounds__(1024) msm_bin_scan_kernel(const uint32_t* __restrict__ bin_counts, uint32_t* __restrict__ bin_offsets,
                                                            uint32_t* __restrict__ chunk_first, uint32_t nbins) {
    __shared__ uint32_t wsum_p[32], wsum_c[32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    uint32_t pad[4], ch[4];
    uint32_t tot_p = 0, tot_c = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t idx = 4u * (uint32_t)t + (uint32_t)j;
        const uint32_t cnt = idx < nbins ? bin_counts[idx] : 0u;
        pad[j] = (cnt + 3u) & ~3u;
        ch[j] = (cnt + (uint32_t)kChunk - 1u) / (uint32_t)kChunk;
        tot_p += pad[j];
        tot_c += ch[j];
    }
    const uint32_t inc_p = warp_inclusive_scan(tot_p, lane), inc_c = warp_inclusive_scan(tot_c, lane);
    if (lane == 31) { wsum_p[warp] = inc_p; wsum_c[warp] = inc_c; }
    __syncthreads();
    if (warp == 0) {
        wsum_p[lane] = warp_inclusive_scan(wsum_p[lane], lane);
        wsum_c[lane] = warp_inclusive_scan(wsum_c[lane], lane);
    }
    __syncthreads();
    uint32_t run_p = (warp ? wsum_p[warp - 1] : 0u) + inc_p - tot_p;
    uint32_t run_c = (warp ? wsum_c[warp - 1] : 0u) + inc_c - tot_c;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t idx = 4u * (uint32_t)t + (uint32_t)j;
        if (idx < nbins) { bin_offsets[idx] = run_p; chunk_first[idx] = run_c; }
        run_p += pad[j];
        run_c += ch[j];
    }
    if (t == 1023) { bin_offsets[nbins] = run_p; chunk_first[nbins] = run_c; }
}

__global__ void __launch_bounds__(kBinThreads) msm_bin_scatter_kernel(const uint4* __restrict__ scalars, const uint32_t* __restrict__ bin_offsets,
                                                                      uint32_t* __restrict__ bin_cursor, uint32_t* __restrict__ part,
                                                                      size_t n, int c, int windows, int groups, uint32_t nb, uint32_t nbins, int table,
                                                                      uint32_t stripe) {
    __shared__ uint32_t s_cnt[kMaxBins];
    __shared__ uint32_t s_base[kMaxBins];
    for (size_t base = (size_t)blockIdx.x * stripe; base < n; base += (size_t)gridDim.x * stripe) {
        for (uint32_t b = threadIdx.x; b < nbins; b += kBinThreads) s_cnt[b] = 0;
        __syncthreads();
        const size_t i = base + threadIdx.x;
        Fr s = Fr::zero();
        if (threadIdx.x < stripe && i < n) s = Fr::load_nc(scalars + 2 * i).from_mont();
        // one counting round: the rank of every entry inside its (stripe, bin) stays in registers, two 16-bit ranks per
        // word (a stripe has at most 1024 * 32 entries)
        uint32_t rk[kMaxWindows / 2];
        {
            DigitStream ds(s);
#pragma unroll
            for (int w = 0; w < kMaxWindows; w++) {
                if (w < windows) {
                    const uint32_t mag = ds.next(c) & 0x7fffffffu;
                    const uint32_t id = (groups > 1 ? (uint32_t)w * nb : 0u) + (mag - 1u);
                    const uint32_t r = smem_rank(s_cnt, id >> kFineBits, mag != 0);
                    if (w & 1) rk[w >> 1] |= r << 16; else rk[w >> 1] = r;
                }
            }
        }
        __syncthreads();
        for (uint32_t b = threadIdx.x; b < nbins; b += kBinThreads) {
            const uint32_t v = s_cnt[b];
            s_base[b] = v ? bin_offsets[b] + atomicAdd(&bin_cursor[b], v) : 0u;
        }
        __syncthreads();
        {
            DigitStream ds(s);
#pragma unroll
            for (int w = 0; w < kMaxWindows; w++) {
                if (w < windows) {
                    const uint32_t d = ds.next(c);
                    const uint32_t mag = d & 0x7fffffffu;
                    if (mag) {
                        const uint32_t id = (groups > 1 ? (uint32_t)w * nb : 0u) + (mag - 1u);
                        const uint32_t r = (rk[w >> 1] >> ((w & 1) * 16)) & 0xffffu;
                        const uint32_t pidx = table ? (uint32_t)((size_t)w * n + i) : (uint32_t)i;   // table[w*n + i] = 2^(c w) P_i
                        part[s_base[id >> kFineBits] + r] = ((id & (uint32_t)(kFine - 1)) << 25) | ((d >> 31) << 24) | pidx;
                    }
                }
            }
        }
        __syncthreads();
    }
}

// (Ranks from match.any + warp-private counters instead of returning shared-memory atomics were built and measured:
// bin_scatter 78 -> 133 us, fine_scatter 48 -> 80 us at 2^19 uniform scalars - MATCH.ANY costs more than the atomic it
// replaces.  Removed.)
// chunk j -> (bin, first entry, length): bins with no chunk share chunk_first with their successor, so the bin is the
// LAST one whose chunk_first is <= j
__device__ __forceinline__ void chunk_range(const uint32_t* __restrict__ bin_offsets, const uint32_t* __restrict__ bin_counts,
                                            const uint32_t* __restrict__ chunk_first, uint32_t nbins, uint32_t j,
                                            uint32_t* bin, uint32_t* start, uint32_t* len) {
    uint32_t lo = 0, hi = nbins;   // chunk_first[lo] <= j < chunk_first[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (chunk_first[mid] <= j) lo = mid; else hi = mid;
    }
    const uint32_t piece = j - chunk_first[lo];
    const uint32_t left = bin_counts[lo] - piece * (uint32_t)kChunk;
    *bin = lo;
    *start = bin_offsets[lo] + piece * (uint32_t)kChunk;
    *len = left < (uint32_t)kChunk ? left : (uint32_t)kChunk;
}

__global__ void __launch_bounds__(kChunkThreads) msm_fine_count_kernel(const uint32_t* __restrict__ part, const uint32_t* __restrict__ bin_offsets,
                                                                       const uint32_t* __restrict__ bin_counts, const uint32_t* __restrict__ chunk_first,
                                                                       uint32_t nbins, uint32_t* __restrict__ counts, uint32_t* __restrict__ chunk_base) {
    __shared__ __align__(128) uint32_t s_e[kChunk];
    __shared__ uint32_t s_cnt[kFine];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t j = blockIdx.x;
    if (j >= chunk_first[nbins]) return;
    uint32_t bin, start, len;
    chunk_range(bin_offsets, bin_counts, chunk_first, nbins, j, &bin, &start, &len);
    if (threadIdx.x < kFine) s_cnt[threadIdx.x] = 0;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        bulk_load(s_e, part + start, ((len + 3u) & ~3u) * 4u, &bar);   // bins are padded to four entries: a legal bulk copy
    }
    __syncthreads();
    mbar_wait(&bar, 0);
#pragma unroll 1
    for (uint32_t k = threadIdx.x; k < (uint32_t)kChunk; k += kChunkThreads) {
        const bool valid = k < len;
        const uint32_t e = valid ? s_e[k] : 0u;
        smem_rank(s_cnt, e >> 25, valid);
    }
    __syncthreads();
    if (threadIdx.x < kFine) {
        const uint32_t v = s_cnt[threadIdx.x];
        chunk_base[(size_t)j * kFine + threadIdx.x] = v ? atomicAdd(&counts[(size_t)bin * kFine + threadIdx.x], v) : 0u;
    }
}

__global__ void __launch_bounds__(kChunkThreads) msm_fine_scatter_kernel(const uint32_t* __restrict__ part, const uint32_t* __restrict__ bin_offsets,
                                                                         const uint32_t* __restrict__ bin_counts, const uint32_t* __restrict__ chunk_first,
                                                                         uint32_t nbins, const uint32_t* __restrict__ offsets,
                                                                         const uint32_t* __restrict__ chunk_base, uint32_t* __restrict__ sorted) {
    __shared__ __align__(128) uint32_t s_e[kChunk];
    __shared__ uint32_t s_cur[kFine];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t j = blockIdx.x;
    if (j >= chunk_first[nbins]) return;
    uint32_t bin, start, len;
    chunk_range(bin_offsets, bin_counts, chunk_first, nbins, j, &bin, &start, &len);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        bulk_load(s_e, part + start, ((len + 3u) & ~3u) * 4u, &bar);
    }
    if (threadIdx.x < kFine) s_cur[threadIdx.x] = offsets[(size_t)bin * kFine + threadIdx.x] + chunk_base[(size_t)j * kFine + threadIdx.x];
    __syncthreads();
    mbar_wait(&bar, 0);
#pragma unroll 1
    for (uint32_t k = threadIdx.x; k < (uint32_t)kChunk; k += kChunkThreads) {
        const bool valid = k < len;
        const uint32_t e = valid ? s_e[k] : 0u;
        const uint32_t pos = smem_rank(s_cur, e >> 25, valid);
        if (valid) sorted[pos] = (((e >> 24) & 1u) << 31) | (e & 0x00ffffffu);
    }
}

// ---- accumulate: one thread per run of plan->run consecutive sorted entries ----------------------------
__global__ void __launch_bounds__(kAccThreads, ZKW_MSM_CTAS_PER_SM)
msm_accumulate_kernel(const uint4* __restrict__ points, const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ offsets,
                      const RunPlan* __restrict__ plan, uint4* __restrict__ partials, uint32_t total_buckets) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t run = plan->run, entries = plan->entries;
    if (t >= plan->threads) return;
    const uint32_t begin = t * run;
    const uint32_t end = begin + run < entries ? begin + run : entries;
    // bucket of the first entry: last b with offsets[b] <= begin (empty buckets share an offset with the next)
    uint32_t lo = 0, hi = total_buckets;  // invariant: offsets[lo] <= begin < offsets[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= begin) lo = mid; else hi = mid;
    }
    uint32_t b = lo;
    uint32_t bend = offsets[b + 1];
    uint32_t e = sorted[begin];
    G1Affine cur = G1Affine::load_nc(points + 4 * (size_t)(e & 0x7fffffffu));
    G1Xyzz acc = G1Xyzz::identity();
    for (uint32_t k = begin; k < end; k++) {
        const bool neg = e >> 31;
        G1Affine nxt = cur;
        if (k + 1 < end) {
            e = sorted[k + 1];
            nxt = G1Affine::load_nc(points + 4 * (size_t)(e & 0x7fffffffu));
        }
        if (k == bend) {   // entry k opens the next non-empty bucket: emit the finished one
            acc.normalize();
            acc.store(partials + 8 * ((size_t)t + b));
            acc = G1Xyzz::identity();
            do { b++; bend = offsets[b + 1]; } while (bend <= k);
        }
#ifdef ZKW_MSM_NO_LAZY
        acc.add_mixed(cur, neg);
#else
        acc.add_mixed_lazy(cur, neg);   // running point kept in [0, 2p): no conditional subtraction after the products
#endif
        cur = nxt;
    }
    acc.normalize();
    acc.store(partials + 8 * ((size_t)t + b));
}

__device__ __forceinline__ G1Xyzz shfl_xor_point(const G1Xyzz& p, int lane_mask, unsigned member_mask) {
    G1Xyzz r;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r.x.l[i] = __shfl_xor_sync(member_mask, p.x.l[i], lane_mask);
        r.y.l[i] = __shfl_xor_sync(member_mask, p.y.l[i], lane_mask);
        r.zz.l[i] = __shfl_xor_sync(member_mask, p.zz.l[i], lane_mask);
        r.zzz.l[i] = __shfl_xor_sync(member_mask, p.zzz.l[i], lane_mask);
    }
    return r;
}

// ---- combine: bucket b's partials live in slots t + b, t = first run .. last run touching the bucket ----
__device__ __forceinline__ bool bucket_partials(const uint32_t* __restrict__ offsets, uint32_t run, uint32_t b, size_t* first, uint32_t* count) {
    const uint32_t o0 = offsets[b], o1 = offsets[b + 1];
    if (o1 == o0) { *count = 0; *first = 0; return false; }
    const uint32_t t0 = o0 / run, t1 = (o1 - 1) / run;
    *first = (size_t)t0 + b;
    *count = t1 - t0 + 1;
    return true;
}

// one thread per bucket: folds up to kLight partials (the common case); heavier buckets are queued for the CTA kernel.
// (64-thread CTAs, so that the 2^16 buckets of a k = 19 MSM are one resident wave at 140 registers, measured: no change, 47 us.)
constexpr int kLightThreads = 128;
__global__ void __launch_bounds__(kLightThreads) msm_combine_light_kernel(const uint4* __restrict__ partials, const uint32_t* __restrict__ offsets,
                                                                RunPlan* __restrict__ plan, uint32_t* __restrict__ heavy_list,
                                                                uint4* __restrict__ buckets, uint32_t total_buckets) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= total_buckets) return;
    size_t first;
    uint32_t count;
    bucket_partials(offsets, plan->run, b, &first, &count);
    if (count > (uint32_t)kMedium) { heavy_list[atomicAdd(&plan->heavy, 1u)] = b; return; }
    if (count > (uint32_t)kLight) { heavy_list[total_buckets - 1u - atomicAdd(&plan->medium, 1u)] = b; return; }   // medium queue grows down from the end of the list
    G1Xyzz acc = G1Xyzz::identity();   // all-zero = identity (ZZ = 0): what an empty bucket stores
    if (count) acc = G1Xyzz::load(partials + 8 * first);
    for (uint32_t i = 1; i < count; i++) {
        G1Xyzz p = G1Xyzz::load(partials + 8 * (first + i));
        acc.add(p);
    }
    acc.store(buckets + 8 * (size_t)b);
}

// Queued buckets (cut into more than kLight runs - skewed scalars: zeros, bits, tiny digits of sorted lookup columns, long
// constant stretches of a grand product), grid-stride over (bucket, segment) items.  A bucket with fewer than 2 * kHeavySeg
// partials is one item: strided loop, shuffle tree, then the eight warp sums through shared memory - all sharing one instance
// of the point addition (runtime-counted loop whose operand comes from memory, from a shuffle, or from shared memory).  A
// heavier bucket (a permuted lookup column leaves three buckets with 8192 partials each) is cut into up to kHeavySplit
// segments, one CTA each, and the CTA that finishes last (a counter per queued bucket) folds the segment sums: the dependent
// chain of additions shrinks from count / 256 + 8 to count / 2048 + 8 + 4.
constexpr int kHeavyThreads = 256;
constexpr int kHeavySplit = 8;
constexpr uint32_t kHeavySeg = 512;    // partials per segment below which a bucket is not split further
__global__ void __launch_bounds__(kHeavyThreads) msm_combine_heavy_kernel(const uint4* __restrict__ partials, const uint32_t* __restrict__ offsets,
                                                                          const RunPlan* __restrict__ plan, const uint32_t* __restrict__ heavy_list,
                                                                          uint4* __restrict__ buckets, uint4* __restrict__ seg_sums,
                                                                          uint32_t* __restrict__ seg_done) {
    __shared__ uint4 sh[(kHeavyThreads / 32) * 8];
    __shared__ uint32_t s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nitems = plan->heavy * (uint32_t)kHeavySplit;
#pragma unroll 1
    for (uint32_t item = blockIdx.x; item < nitems; item += gridDim.x) {
    const uint32_t h = item / (uint32_t)kHeavySplit, seg = item % (uint32_t)kHeavySplit;
    const uint32_t b = heavy_list[h];
    size_t first;
    uint32_t count;
    bucket_partials(offsets, plan->run, b, &first, &count);
    uint32_t nseg = count / kHeavySeg;
    nseg = nseg < 1u ? 1u : (nseg > (uint32_t)kHeavySplit ? (uint32_t)kHeavySplit : nseg);
    if (seg >= nseg) continue;                                   // CTA-uniform
    const uint32_t per = (count + nseg - 1) / nseg;
    const uint32_t lo = seg * per, hi = lo + per < count ? lo + per : count;
    // phase 0: this segment's partials; phase 1 (last CTA of a split bucket only): the nseg segment sums
    for (int phase = 0; phase < 2; phase++) {
    const uint4* src = phase == 0 ? partials + 8 * (first + lo) : seg_sums + 8 * ((size_t)h * kHeavySplit);
    const uint32_t cnt = phase == 0 ? hi - lo : nseg;
    G1Xyzz acc = G1Xyzz::identity();
    const int nload = (int)((cnt + kHeavyThreads - 1) / kHeavyThreads);
    const int total_it = nload + 5 + 3;
#pragma unroll 1
    for (int it = 0; it < total_it; it++) {
        if (it == nload + 5) {   // every lane of a warp holds the warp's sum: hand the 8 sums to warp 0
            if (lane == 0) acc.store(sh + 8 * warp);
            __syncthreads();
            acc = (warp == 0 && lane < kHeavyThreads / 32) ? G1Xyzz::load(sh + 8 * lane) : G1Xyzz::identity();
        }
        G1Xyzz o = G1Xyzz::identity();
        if (it < nload) {
            const uint32_t i = (uint32_t)it * kHeavyThreads + threadIdx.x;
            if (i < cnt) o = phase == 0 ? G1Xyzz::load(src + 8 * (size_t)i) : G1Xyzz::load_cg(src + 8 * (size_t)i);
        } else if (it < nload + 5) {
            o = shfl_xor_point(acc, 16 >> (it - nload), 0xffffffffu);
        } else {
            o = shfl_xor_point(acc, 4 >> (it - nload - 5), 0xffffffffu);
        }
        acc.add(o);
    }
    __syncthreads();   // sh is reused by the next phase / item
    if (nseg == 1 || phase == 1) {
        if (threadIdx.x == 0) acc.store(buckets + 8 * (size_t)b);
        break;
    }
    // split bucket: publish this segment's sum; whoever arrives last folds them
    if (threadIdx.x == 0) {
        acc.store(seg_sums + 8 * ((size_t)h * kHeavySplit + seg));
        __threadfence();
        const uint32_t prev = atomicAdd(&seg_done[h], 1u);
        s_last = prev == nseg - 1 ? 1u : 0u;                       // (seg_done is cleared with the bucket counts, per MSM)
    }
    __syncthreads();
    if (!s_last) break;                                          // CTA-uniform
    }
    }
}

// ---- bucket reduction: sum_b b * B_b by rows and columns ------------------------------------------------
// Bucket index i = b - 1 = hi * 2^lb + lo.  With row sums R_hi = sum_lo B[hi][lo] and column sums
// C_lo = sum_hi B[hi][lo]:   sum_b b * B_b = 2^lb * sum_hi hi * R_hi + sum_lo (lo + 1) * C_lo.
// Kernel 1 forms the 2^(c-1-lb) + 2^lb row and column sums (one warp each: strided loads, shuffle tree);
// kernel 2 bit-slices the two short weighted sums (one warp per weight bit).  About 2 * 2^(c-1) additions
// instead of the c * 2^(c-2) of bit-slicing all buckets, and two short dependent chains.
// The last 2c doublings/additions run on the host in microseconds (msm_host_tail).
// One warp per row / column / weight bit: every lane sums its strided share sequentially, then a five-level
// shuffle tree.  (A wider tree would shorten the dependent chain but every level of it costs a full
// addition on all 32 lanes; the sequential share keeps the arithmetic close to the useful 2 * 2^(c-1) additions.)
// The strided loop and the tree share ONE instance of the ~2300-instruction point addition (a runtime-counted
// loop whose operand comes from memory or from a shuffle): these warps run alone on their SM sub-partition,
// so a second copy of the addition per tree level would be paid in instruction-cache misses.
struct StridedSum {
    const uint4* base;   // element e lives at base + 8 * (first + e * stride)
    size_t first, stride;
    uint32_t count;      // elements in the whole sequence
    uint32_t filter_bit; // 32: take every element; else take element e iff ((e + filter_bias) >> filter_bit) & 1
    uint32_t filter_bias;
};

__device__ __forceinline__ G1Xyzz warp_strided_sum(const StridedSum& q, int lane) {
    G1Xyzz acc = G1Xyzz::identity();
    const int nload = (int)((q.count + 31) / 32);
#pragma unroll 1
    for (int it = 0; it < nload + 5; it++) {
        G1Xyzz o = G1Xyzz::identity();
        if (it < nload) {
            const uint32_t e = (uint32_t)it * 32u + (uint32_t)lane;
            const bool take = e < q.count && (q.filter_bit >= 32 || (((e + q.filter_bias) >> q.filter_bit) & 1u));
            if (take) o = G1Xyzz::load(q.base + 8 * (q.first + (size_t)e * q.stride));
        } else {
            o = shfl_xor_point(acc, 16 >> (it - nload), 0xffffffffu);
        }
        acc.add(o);
    }
    return acc;  // every lane holds the sum
}

// buckets with kLight < partials <= kMedium (the real witness column has ~400 of them: the top digits of 88-bit limbs,
// carries): one warp each - at most four strided loads per lane and the five-level tree - instead of a 256-thread CTA
// and its eight levels; 4 x SMs CTAs of two warps take the whole queue in one round.
__global__ void __launch_bounds__(kReduceThreads) msm_combine_medium_kernel(const uint4* __restrict__ partials, const uint32_t* __restrict__ offsets,
                                                                            const RunPlan* __restrict__ plan, const uint32_t* __restrict__ heavy_list,
                                                                            uint4* __restrict__ buckets, uint32_t total_buckets) {
    const int lane = threadIdx.x & 31;
    const uint32_t nwarps = gridDim.x * (kReduceThreads / 32), nmedium = plan->medium;
#pragma unroll 1
    for (uint32_t m = blockIdx.x * (kReduceThreads / 32) + (threadIdx.x >> 5); m < nmedium; m += nwarps) {   // warp-uniform
        const uint32_t b = heavy_list[total_buckets - 1u - m];
        StridedSum q;
        q.base = partials;
        q.stride = 1;
        q.filter_bit = 32; q.filter_bias = 0;
        bucket_partials(offsets, plan->run, b, &q.first, &q.count);
        const G1Xyzz acc = warp_strided_sum(q, lane);
        if (lane == 0) acc.store(buckets + 8 * (size_t)b);
    }
}

// The same sum by a whole CTA of THREADS: strided loads, the five-level shuffle tree, the warp sums through shared memory and
// log2(THREADS / 32) more shuffle levels in warp 0 - count / THREADS + 5 + log2(THREADS / 32) dependent additions instead of
// count / 32 + 5 (256 elements, 128 threads: 9 instead of 13), still ONE instance of the addition.  Thread 0 returns the sum.
template <int THREADS>
__device__ __forceinline__ G1Xyzz cta_strided_sum(const StridedSum& q, uint4* sh /* (THREADS / 32) * 8 uint4 */) {
    constexpr int kW = THREADS / 32;
    constexpr int kLw = kW == 1 ? 0 : (kW == 2 ? 1 : (kW == 4 ? 2 : 3));
    static_assert(kW == 1 || kW == 2 || kW == 4 || kW == 8, "cta_strided_sum: 32, 64, 128 or 256 threads");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    G1Xyzz acc = G1Xyzz::identity();
    const int nload = (int)((q.count + THREADS - 1) / THREADS);
#pragma unroll 1
    for (int it = 0; it < nload + 5 + kLw; it++) {
        if (kW > 1 && it == nload + 5) {   // every lane of a warp holds the warp's sum: hand the sums to warp 0
            if (lane == 0) acc.store(sh + 8 * warp);
            __syncthreads();
            acc = (warp == 0 && lane < kW) ? G1Xyzz::load(sh + 8 * lane) : G1Xyzz::identity();
        }
        G1Xyzz o = G1Xyzz::identity();
        if (it < nload) {
            const uint32_t e = (uint32_t)it * THREADS + threadIdx.x;
            const bool take = e < q.count && (q.filter_bit >= 32 || (((e + q.filter_bias) >> q.filter_bit) & 1u));
            if (take) o = G1Xyzz::load(q.base + 8 * (q.first + (size_t)e * q.stride));
        } else if (it < nload + 5) {
            o = shfl_xor_point(acc, 16 >> (it - nload), 0xffffffffu);
        } else {
            o = shfl_xor_point(acc, (kW / 2) >> (it - nload - 5), 0xffffffffu);
        }
        acc.add(o);
    }
    return acc;
}

// THREADS per row / column / weight bit.  These kernels end every MSM and with it every transcript round, at 2-4 % occupancy;
// what they cost is the dependent chain of additions, until there are enough warps to make them throughput-bound (see the
// launch site).
constexpr int kRowThreads = 128;
template <int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 128 ? 4 : 1) msm_rowcol_kernel(const uint4* __restrict__ buckets, uint4* __restrict__ rc,
                                                                                     uint32_t nb, int lb) {
    __shared__ uint4 sh[(THREADS / 32) * 8];
    const uint32_t ncols = 1u << lb, nrows = nb >> lb;
    const uint32_t g = blockIdx.y;
    const uint32_t id = blockIdx.x;
    StridedSum q;
    q.base = buckets + 8 * (size_t)g * nb;
    q.filter_bit = 32; q.filter_bias = 0;
    if (id < nrows) { q.first = (size_t)id << lb; q.stride = 1; q.count = ncols; }
    else { q.first = id - nrows; q.stride = ncols; q.count = nrows; }
    const G1Xyzz acc = cta_strided_sum<THREADS>(q, sh);
    if (threadIdx.x == 0) acc.store(rc + 8 * ((size_t)g * (nrows + ncols) + id));
}

// out[g][j]: j < rbits -> sum of rows whose index has bit j; else sum of columns whose (index + 1) has bit j - rbits
template <int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 128 ? 4 : 1) msm_weighted_kernel(const uint4* __restrict__ rc, uint4* __restrict__ out,
                                                                                       uint32_t nb, int lb, int c) {
    __shared__ uint4 sh[(THREADS / 32) * 8];
    const uint32_t ncols = 1u << lb, nrows = nb >> lb;
    const int rbits = c - 1 - lb;
    const int j = blockIdx.x, g = blockIdx.y;
    StridedSum q;
    q.base = rc + 8 * (size_t)g * (nrows + ncols);
    q.stride = 1;
    if (j < rbits) { q.first = 0; q.count = nrows; q.filter_bit = (uint32_t)j; q.filter_bias = 0; }
    else { q.first = nrows; q.count = ncols; q.filter_bit = (uint32_t)(j - rbits); q.filter_bias = 1; }
    const G1Xyzz acc = cta_strided_sum<THREADS>(q, sh);
    if (threadIdx.x == 0) acc.store(out + 8 * ((size_t)g * c + j));
}

// ---- window tables for a resident basis: table[w*n + i] = 2^(c w) P_i ----------------------------------
// One thread per point: doubles in XYZZ, then normalises its W-1 multiples with one shared inversion.
template <int MAXW>
__global__ void __launch_bounds__(128) msm_table_kernel(const uint4* __restrict__ points, uint4* __restrict__ table, size_t n, int c,
                                                        int windows) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1Affine p = G1Affine::load_nc(points + 4 * i);
    p.store(table + 4 * i);
    if (p.is_identity()) {
        for (int w = 1; w < windows; w++) p.store(table + 4 * ((size_t)w * n + i));
        return;
    }
    G1Xyzz cur = G1Xyzz::from_affine(p);
    Fq xs[MAXW], ys[MAXW], zz[MAXW], zzz[MAXW], pref[MAXW];
    Fq run = Fq::one();
    for (int w = 1; w < windows; w++) {
        for (int d = 0; d < c; d++) cur = cur.dbl();
        xs[w] = cur.x; ys[w] = cur.y; zz[w] = cur.zz; zzz[w] = cur.zzz;
        pref[w] = run;                      // product of t_1..t_{w-1}, t = zz*zzz
        run = run * (cur.zz * cur.zzz);
    }
    Fq inv = fp_inv_bingcd(run);
    for (int w = windows - 1; w >= 1; w--) {
        Fq tinv = inv * pref[w];            // 1/(zz_w * zzz_w)
        inv = inv * (zz[w] * zzz[w]);
        G1Affine q;
        q.x = xs[w] * (zzz[w] * tinv);      // X / ZZ
        q.y = ys[w] * (zz[w] * tinv);       // Y / ZZZ
        q.store(table + 4 * ((size_t)w * n + i));
    }
}

// ---- Jacobian -> affine, one thread per point (C::Curve::batch_normalize; small m) ---------------------
__global__ void g1_normalize_kernel(const uint4* __restrict__ xyz, uint4* __restrict__ xy, size_t m) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    Fq x = Fq::load(xyz + 6 * i), y = Fq::load(xyz + 6 * i + 2), z = Fq::load(xyz + 6 * i + 4);
    G1Affine a;
    if (z.is_zero()) { a.x = Fq::zero(); a.y = Fq::zero(); }
    else {
        Fq zi = fp_inv_bingcd(z), zi2 = zi.sqr();
        a.x = x * zi2;
        a.y = y * (zi2 * zi);
    }
    a.store(xy + 4 * i);
}

int g1_batch_normalize_dev(zkw_ctx* ctx, const uint64_t* xyz_dev, size_t m, uint64_t* out_xy_dev) {
    if (m == 0) return ZKW_OK;
    { ProfScope ps_(ctx, "g1_normalize_kernel"); g1_normalize_kernel<<<(unsigned)((m + 63) / 64), 64, 0, ctx->stream>>>((const uint4*)xyz_dev, (uint4*)out_xy_dev, m); }
    ZKW_LAUNCHED(ctx);
    return ZKW_OK;
}

static int pick_window_bits(const zkw_ctx* ctx, size_t n) {
    if (ctx->msm_window_bits > 0) return ctx->msm_window_bits;
    // What matters is the window COUNT ceil(255 / c): the narrowest c of each count wins (fewest buckets for the same
    // number of additions).  Measured on B200 (tools/window_sweep_large.py, ms per uniform MSM):
    //   2^20: c=16 3.56, c=17 3.30, c=20 3.76      2^21: c=16 6.82, c=17 6.23, c=20 6.39
    //   2^22: c=17 12.2, c=20 11.6, c=22 14.0      2^24: c=17 48.7, c=20 43.1, c=22 46.1
    //   2^19: c=16 2.02, c=17 2.01 (a tie; 16 keeps the tables' 16 windows)
    // With the binned entry sort and the cheaper bucket combine (round 2) the 2^19 tie went to c = 17: 15 windows instead of
    // 16 (-6 % additions) against twice the buckets: MSM 1.709 -> 1.634 ms, k = 19 proof 25.23 -> 24.71 ms (c = 19: 1.865 / 26.12).
    // Below 2^17 points the bucket reduction's dependent chain is a third of the MSM and fewer buckets win: c = 15 (17 windows,
    // 2^14 buckets) against 16 - 2^16 points 0.458 / 0.509 ms, 2^15 points 0.339 / 0.365 ms; at 2^17 points 0.672 / 0.653 ms alone
    // (the k = 17 proof, resident: 11.13 against 11.40 ms in one run, its end-to-end time unchanged) - 16 stays there.  (Narrower
    // windows put more than kLight partials into EVERY bucket - the run length cannot go below kMinRun - and the whole bucket
    // set lands in the warp-per-bucket queue: c = 14 at 2^17: 1.12 ms.)
    if (n >= (1u << 22)) return 20;
    if (n >= (1u << 19)) return 17;
    if (n >= (1u << 17)) return 16;
    return n >= (1u << 13) ? 15 : 8;
}

int msm_window_bits(const zkw_ctx* ctx, size_t n) { return pick_window_bits(ctx, n); }

static void make_plan(MsmPlan& p, size_t n, int c, bool table) {
    p.c = c;
    p.windows = (255 + c - 1) / c;
    p.groups = table ? 1 : p.windows;
    p.nb = 1u << (c - 1);
    p.n = n;
}

void msm_free_basis(MsmBasis& b) {
    if (b.points) cudaFree(b.points);
    if (b.table) cudaFree(b.table);
    b = MsmBasis();
}

int msm_prepare_basis(zkw_ctx* ctx, MsmBasis& b) {
    if (!ctx->msm_precompute || b.n == 0) return ZKW_OK;
    const int c = pick_window_bits(ctx, b.n);
    MsmPlan p;
    make_plan(p, b.n, c, true);
    if ((size_t)p.windows * b.n >= (1ull << 31)) return ZKW_OK;  // entry ids carry the sign in bit 31
    if (p.windows > 32) return ZKW_OK;
    if (b.table) { cudaFree(b.table); b.table = nullptr; }
    cudaError_t e = cudaMalloc((void**)&b.table, (size_t)p.windows * b.n * 64);
    if (e != cudaSuccess) { b.table = nullptr; cudaGetLastError(); return ZKW_OK; }  // fall back to per-window groups
    { ProfScope ps_(ctx, "msm_table_kernel"); msm_table_kernel<32><<<(unsigned)((b.n + 127) / 128), 128, 0, ctx->stream>>>((const uint4*)b.points, (uint4*)b.table, b.n, c, p.windows); }
    ZKW_LAUNCHED(ctx);
    b.c = c;
    b.windows = p.windows;
    return ZKW_OK;
}

static inline int rowcol_lb(int c) { return c / 2; }  // column bits of the bucket index split

// host-side tail: sum_g 2^(c g) * (2^lb * sum_j 2^j SR[g][j] + sum_j 2^j SC[g][j]), then affine normalisation
static void msm_host_tail(const uint32_t* s_xyzz /* [G][c][32] */, int groups, int c, uint64_t out_xyz[12]) {
    const int lb = rowcol_lb(c), rbits = c - 1 - lb, cbits = lb + 1;
    auto at = [&](int g, int j) { G1Xyzz s; memcpy(&s, s_xyzz + 32 * ((size_t)g * c + j), 128); return s; };
    G1Xyzz acc = G1Xyzz::identity();
    for (int g = groups - 1; g >= 0; g--) {
        for (int d = 0; d < c && g != groups - 1; d++) acc = acc.dbl();
        G1Xyzz rows = G1Xyzz::identity();
        for (int j = rbits - 1; j >= 0; j--) { rows = rows.dbl(); rows.add(at(g, j)); }
        for (int d = 0; d < lb; d++) rows = rows.dbl();
        G1Xyzz cols = G1Xyzz::identity();
        for (int j = cbits - 1; j >= 0; j--) { cols = cols.dbl(); cols.add(at(g, rbits + j)); }
        acc.add(rows);
        acc.add(cols);
    }
    Fq x = Fq::zero(), y = Fq::one(), z = Fq::zero();
    if (!acc.is_identity()) {
        // x = X/ZZ, y = Y/ZZZ with one inversion of ZZ*ZZZ
        Fq tinv = fp_inv_bingcd(acc.zz * acc.zzz);
        x = acc.x * (acc.zzz * tinv);
        y = acc.y * (acc.zz * tinv);
        z = Fq::one();
    }
    memcpy(out_xyz, x.l, 32);
    memcpy(out_xyz + 4, y.l, 32);
    memcpy(out_xyz + 8, z.l, 32);
}

}  // namespace zkw

// Sum of m Jacobian points (x, y, z Montgomery; z = 0 is the identity) on the HOST, normalised to (x, y, 1): the fold of
// the per-GPU partial results of a split MSM - m - 1 point additions and one inversion, microseconds.
extern "C" int zkw_g1_sum(const uint64_t* xyz, size_t m, uint64_t out_xyz[12]) {
    using namespace zkw;
    if ((!xyz && m) || !out_xyz) return ZKW_ERR_INVALID;
    G1Xyzz acc = G1Xyzz::identity();
    for (size_t i = 0; i < m; i++) {
        Fq x, y, z;
        memcpy(x.l, xyz + 12 * i, 32); memcpy(y.l, xyz + 12 * i + 4, 32); memcpy(z.l, xyz + 12 * i + 8, 32);
        if (z.is_zero()) continue;
        G1Xyzz p;
        p.x = x; p.y = y; p.zz = z * z; p.zzz = p.zz * z;
        acc.add(p);
    }
    Fq x = Fq::zero(), y = Fq::one(), z = Fq::zero();
    if (!acc.is_identity()) {
        Fq tinv = fp_inv_bingcd(acc.zz * acc.zzz);
        x = acc.x * (acc.zzz * tinv);
        y = acc.y * (acc.zz * tinv);
        z = Fq::one();
    }
    memcpy(out_xyz, x.l, 32); memcpy(out_xyz + 4, y.l, 32); memcpy(out_xyz + 8, z.l, 32);
    return ZKW_OK;
}

namespace zkw {

static int lane_init(zkw_ctx* ctx, int lane) {
    if (!ctx->fork_event) ZKW_CUDA(ctx, cudaEventCreateWithFlags(&ctx->fork_event, cudaEventDisableTiming));
    if (lane == 0) { ctx->lane_stream[0] = ctx->stream; }
    else if (!ctx->lane_stream[lane]) ZKW_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->lane_stream[lane], cudaStreamNonBlocking, stream_priority(0)));
    if (!ctx->lane_done[lane]) ZKW_CUDA(ctx, cudaEventCreateWithFlags(&ctx->lane_done[lane], cudaEventDisableTiming));
    return ZKW_OK;
}

// Enqueue one MSM on a lane's stream (no host synchronisation); results land in the lane's pinned slot.
static int msm_enqueue(zkw_ctx* ctx, int lane, int which_bases, const uint64_t* bases_dev, const uint64_t* scalars_dev, size_t n) {
    ZKW_TRY(lane_init(ctx, lane));
    cudaStream_t st = ctx->lane_stream[lane];
    ctx->lane_groups[lane] = 0;
    if (n == 0) return ZKW_OK;
    const uint64_t* points = bases_dev;
    bool table = false;
    int c = pick_window_bits(ctx, n);
    if (which_bases == ZKW_BASES_G || which_bases == ZKW_BASES_G_LAGRANGE) {
        MsmBasis& b = ctx->bases[which_bases];
        if (!b.points) return ZKW_ERR_STATE;
        if (n > b.n) return ZKW_ERR_INVALID;
        if (b.table && n == b.n) { points = b.table; table = true; c = b.c; }
        else points = b.points;
    } else if (which_bases != ZKW_BASES_CALLER || !bases_dev) {
        return ZKW_ERR_INVALID;
    }
    MsmPlan p;
    make_plan(p, n, c, table);
    if (p.max_entries() >= (1ull << 31)) return ZKW_ERR_INVALID;
    const size_t tb = p.total_buckets();
    const uint32_t max_threads = (uint32_t)ctx->sm_count * ZKW_MSM_CTAS_PER_SM * kAccThreads * ZKW_MSM_WAVES;  // whole resident waves
    const size_t acc_threads = std::min<size_t>(max_threads, (p.max_entries() + kMinRun - 1) / kMinRun);
    const size_t n_partials = acc_threads + tb;   // slot t + b, t < acc_threads, b < tb
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    // binned sort (the default wherever its packed entry fits): bucket id = bin * 128 + fine
    const uint32_t nbins = (uint32_t)(tb >> kFineBits);
    const bool binned = ctx->msm_binned_sort && (tb & (kFine - 1)) == 0 && nbins >= 1 && nbins <= (uint32_t)kMaxBins &&
                        p.max_entries() <= kBinnedMaxEntries && p.windows <= kMaxWindows && p.max_entries() >= (size_t)ctx->msm_binned_min_entries;
    const size_t max_chunks = binned ? (size_t)nbins + p.max_entries() / kChunk + 1 : 0;
    const size_t o_digits = take((p.max_entries() + (binned ? 4 * (size_t)nbins + 4 : 0)) * 4);   // binned: the partitioned entries, bins padded
    const size_t o_sorted = take(p.max_entries() * 4);
    const size_t o_counts = take(tb * 4);
    const size_t o_cursor = take(tb * 4);        // binned: bin_counts | bin_cursor (2 * nbins <= tb words)
    // split heavy buckets: a queued bucket has more than kLight partials, so at most n_partials / kLight of them exist
    const size_t max_heavy = std::min<size_t>(tb, n_partials / kLight + 1);
    const size_t o_seg_done = take(max_heavy * 4);   // cleared with counts and cursor
    const size_t o_offsets = take((tb + 1) * 4);
    const size_t o_bin_offsets = take(((size_t)nbins + 1) * 4);
    const size_t o_chunk_first = take(((size_t)nbins + 1) * 4);
    const size_t o_chunk_base = take(max_chunks * kFine * 4);
    const size_t o_plan = take(sizeof(RunPlan));
    const size_t o_heavy = take(tb * 4);
    const size_t o_seg_sums = take(max_heavy * kHeavySplit * 128);
    const size_t o_partials = take(n_partials * 128);
    const size_t o_buckets = take(tb * 128);
    const int lb = rowcol_lb(c);
    const uint32_t rc_per_group = (p.nb >> lb) + (1u << lb);
    const size_t o_blocks = take((size_t)p.groups * rc_per_group * 128);
    const size_t o_out = take((size_t)p.groups * c * 128);
    DeviceBuffer& wsb = lane == 0 ? ctx->msm_ws : ctx->lane_ws[lane];
    if (wsb.bytes < off) {
        // growing a lane workspace frees memory other lanes' queued work does not touch, but cudaFree
        // synchronises the device anyway; this only happens on the first MSM of a new size
        if (wsb.ptr) { cudaDeviceSynchronize(); cudaFree(wsb.ptr); wsb.ptr = nullptr; wsb.bytes = 0; }
        ZKW_CUDA(ctx, cudaMalloc(&wsb.ptr, off));
        wsb.bytes = off;
    }
    char* ws = (char*)wsb.ptr;
    uint32_t* digits = (uint32_t*)(ws + o_digits);
    uint32_t* sorted = (uint32_t*)(ws + o_sorted);
    uint32_t* counts = (uint32_t*)(ws + o_counts);
    uint32_t* cursor = (uint32_t*)(ws + o_cursor);
    uint32_t* offsets = (uint32_t*)(ws + o_offsets);
    RunPlan* plan = (RunPlan*)(ws + o_plan);
    uint32_t* heavy_list = (uint32_t*)(ws + o_heavy);
    uint4* partials = (uint4*)(ws + o_partials);
    uint4* seg_sums = (uint4*)(ws + o_seg_sums);
    uint32_t* seg_done = (uint32_t*)(ws + o_seg_done);
    uint4* buckets = (uint4*)(ws + o_buckets);
    uint4* blocks = (uint4*)(ws + o_blocks);
    uint4* outs = (uint4*)(ws + o_out);
    ZKW_CUDA(ctx, cudaMemsetAsync(counts, 0, (o_seg_done - o_counts) + max_heavy * 4, st));
    // (partial slots that no run writes - a run boundary coinciding with a bucket boundary, empty buckets - are
    // never read either: a bucket's slot range covers exactly the runs that intersect it)
    if (binned) {
        uint32_t* bin_counts = cursor;
        uint32_t* bin_cursor = cursor + nbins;
        uint32_t* bin_offsets = (uint32_t*)(ws + o_bin_offsets);
        uint32_t* chunk_first = (uint32_t*)(ws + o_chunk_first);
        uint32_t* chunk_base = (uint32_t*)(ws + o_chunk_base);
        // One CTA of 1024 threads is resident per SM (48 registers), so the scalars are cut into stripes such that every SM
        // gets the same number of them: 2^19 scalars over 148 SMs = 3.46 full stripes, i.e. four rounds with a quarter of the
        // SMs idle in the last; 592 stripes of 896 scalars are four full rounds.
        const size_t sms = (size_t)ctx->sm_count;
        const size_t rounds = (n + sms * kBinThreads - 1) / (sms * kBinThreads);
        uint32_t stripe = (uint32_t)((n + sms * rounds - 1) / (sms * rounds));
        stripe = std::min<uint32_t>(kBinThreads, (stripe + 31u) & ~31u);
        const unsigned stripes = (unsigned)std::min<size_t>((n + stripe - 1) / stripe, sms);
        { ProfScope ps_(ctx, "msm_bin_count_kernel", st); msm_bin_count_kernel<<<stripes, kBinThreads, 0, st>>>((const uint4*)scalars_dev, bin_counts, n, c, p.windows, p.groups, p.nb, nbins, stripe); }
        ZKW_LAUNCHED(ctx);
        { ProfScope ps_(ctx, "msm_bin_scan_kernel", st); msm_bin_scan_kernel<<<1, 1024, 0, st>>>(bin_counts, bin_offsets, chunk_first, nbins); }
        ZKW_LAUNCHED(ctx);
        { ProfScope ps_(ctx, "msm_bin_scatter_kernel", st); msm_bin_scatter_kernel<<<stripes, kBinThreads, 0, st>>>((const uint4*)scalars_dev, bin_offsets, bin_cursor, digits, n, c, p.windows, p.groups, p.nb, nbins, table ? 1 : 0, stripe); }
        ZKW_LAUNCHED(ctx);
        { ProfScope ps_(ctx, "msm_fine_count_kernel", st); msm_fine_count_kernel<<<(unsigned)max_chunks, kChunkThreads, 0, st>>>(digits, bin_offsets, bin_counts, chunk_first, nbins, counts, chunk_base); }
        ZKW_LAUNCHED(ctx);
        { ProfScope ps_(ctx, "msm_scan_kernel", st); msm_scan_kernel<<<1, 1024, 0, st>>>(counts, offsets, plan, tb, (uint32_t)acc_threads); }
        ZKW_LAUNCHED(ctx);
        { ProfScope ps_(ctx, "msm_fine_scatter_kernel", st); msm_fine_scatter_kernel<<<(unsigned)max_chunks, kChunkThreads, 0, st>>>(digits, bin_offsets, bin_counts, chunk_first, nbins, offsets, chunk_base, sorted); }
        ZKW_LAUNCHED(ctx);
    } else {
    { ProfScope ps_(ctx, "msm_recode_kernel", st); msm_recode_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>((const uint4*)scalars_dev, digits, counts, n, c, p.windows, p.groups, p.nb); }
    ZKW_LAUNCHED(ctx);
    { ProfScope ps_(ctx, "msm_scan_kernel", st); msm_scan_kernel<<<1, 1024, 0, st>>>(counts, offsets, plan, tb, (uint32_t)acc_threads); }
    ZKW_LAUNCHED(ctx);
    { ProfScope ps_(ctx, "msm_scatter_kernel", st); msm_scatter_kernel<<<(unsigned)((p.max_entries() + 255) / 256), 256, 0, st>>>(digits, offsets, cursor, sorted, n, p.windows, p.groups, p.nb, table ? 1 : 0); }
    ZKW_LAUNCHED(ctx);
    }
    if (kAccSmemReserve > 0 && !ctx->msm_attr_set) {
        ZKW_CUDA(ctx, cudaFuncSetAttribute(msm_accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAccSmemReserve));
        ctx->msm_attr_set = true;
    }
    { ProfScope ps_(ctx, "msm_accumulate_kernel", st); msm_accumulate_kernel<<<(unsigned)((acc_threads + kAccThreads - 1) / kAccThreads), kAccThreads, kAccSmemReserve, st>>>((const uint4*)points, sorted, offsets, plan, partials, (uint32_t)tb); }
    ZKW_LAUNCHED(ctx);
    { ProfScope ps_(ctx, "msm_combine_light_kernel", st); msm_combine_light_kernel<<<(unsigned)((tb + kLightThreads - 1) / kLightThreads), kLightThreads, 0, st>>>(partials, offsets, plan, heavy_list, buckets, (uint32_t)tb); }
    ZKW_LAUNCHED(ctx);
    { ProfScope ps_(ctx, "msm_combine_heavy_kernel", st); msm_combine_heavy_kernel<<<(unsigned)std::min<size_t>(tb, 2 * (size_t)ctx->sm_count), kHeavyThreads, 0, st>>>(partials, offsets, plan, heavy_list, buckets, seg_sums, seg_done); }
    ZKW_LAUNCHED(ctx);
    { ProfScope ps_(ctx, "msm_combine_medium_kernel", st); msm_combine_medium_kernel<<<(unsigned)std::min<size_t>((tb + 1) / 2, 4 * (size_t)ctx->sm_count), kReduceThreads, 0, st>>>(partials, offsets, plan, heavy_list, buckets, (uint32_t)tb); }
    ZKW_LAUNCHED(ctx);
    // Row / column sums: one WARP per sum.  (A CTA of 128 per sum - 9 dependent additions instead of 13 - was measured: 2048
    // warps instead of 512 turn the kernel from latency-bound into throughput-bound, 88 -> 136 us.)  The 17 weighted sums do
    // get a CTA each when there is one bucket set: 81 -> 66 us.
    const bool wide_rows = p.groups == 1;
    { ProfScope ps_(ctx, "msm_rowcol_kernel", st); msm_rowcol_kernel<32><<<dim3(rc_per_group, p.groups), 32, 0, st>>>(buckets, blocks, p.nb, lb); }
    ZKW_LAUNCHED(ctx);
    const size_t out_bytes = (size_t)p.groups * c * 128;
    if (ctx->lane_pinned_bytes[lane] < out_bytes) {
        if (ctx->lane_pinned[lane]) { cudaStreamSynchronize(st); cudaFreeHost(ctx->lane_pinned[lane]); }
        ctx->lane_pinned[lane] = nullptr;
        ctx->lane_pinned_bytes[lane] = 0;
        ZKW_CUDA(ctx, cudaMallocHost(&ctx->lane_pinned[lane], out_bytes));
        ctx->lane_pinned_bytes[lane] = out_bytes;
    }
    // The weighted sums (c x 128 B per group) are stored straight into the lane's page-locked host slot (unified addressing:
    // the kernel writes over PCIe), not copied back: a D2H copy queued behind this lane's kernels sat in the copy queue and
    // held back the H2D copy of the advice columns that the end-to-end path issues while this MSM is still running
    // (measured: the advice arrived 1.1 ms after the random-polynomial MSM had finished instead of under it).
    uint4* outs_host = ctx->msm_zero_copy_out ? (uint4*)ctx->lane_pinned[lane] : outs;
    {
        ProfScope ps_(ctx, "msm_weighted_kernel", st);
        if (wide_rows) msm_weighted_kernel<kRowThreads><<<dim3(c, p.groups), kRowThreads, 0, st>>>(blocks, outs_host, p.nb, lb, c);
        else msm_weighted_kernel<32><<<dim3(c, p.groups), 32, 0, st>>>(blocks, outs_host, p.nb, lb, c);
    }
    ZKW_LAUNCHED(ctx);
    if (!ctx->msm_zero_copy_out) ZKW_CUDA(ctx, cudaMemcpyAsync(ctx->lane_pinned[lane], outs, out_bytes, cudaMemcpyDeviceToHost, st));
    ctx->lane_groups[lane] = p.groups;
    ctx->lane_c[lane] = c;
    return ZKW_OK;
}

static void msm_collect(zkw_ctx* ctx, int lane, uint64_t out_xyz[12]) {
    if (ctx->lane_groups[lane] == 0) {  // empty MSM: identity
        memset(out_xyz, 0, 96);
        Fq one = Fq::one();
        memcpy(out_xyz + 4, one.l, 32);
        return;
    }
    msm_host_tail((const uint32_t*)ctx->lane_pinned[lane], ctx->lane_groups[lane], ctx->lane_c[lane], out_xyz);
}

int msm_run(zkw_ctx* ctx, int which_bases, const uint64_t* bases_dev, const uint64_t* scalars_dev, size_t n,
            uint64_t out_xyz_host[12]) {
    ZKW_TRY(msm_enqueue(ctx, 0, which_bases, bases_dev, scalars_dev, n));
    ZKW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    msm_collect(ctx, 0, out_xyz_host);
    return ZKW_OK;
}

int msm_run_batch(zkw_ctx* ctx, const MsmJob* jobs, int count, uint64_t (*outs)[12]) {
    int done = 0;
    while (done < count) {
        const int m = count - done < zkw_ctx::kMsmLanes ? count - done : zkw_ctx::kMsmLanes;
        ZKW_TRY(lane_init(ctx, 0));
        // side lanes start after everything queued so far on the main stream (their inputs live there)
        ZKW_CUDA(ctx, cudaEventRecord(ctx->fork_event, ctx->stream));
        for (int i = 0; i < m; i++) {
            ZKW_TRY(lane_init(ctx, i));
            if (i) ZKW_CUDA(ctx, cudaStreamWaitEvent(ctx->lane_stream[i], ctx->fork_event, 0));
            const MsmJob& j = jobs[done + i];
            ZKW_TRY(msm_enqueue(ctx, i, j.which_bases, j.bases_dev, j.scalars_dev, j.n));
            if (i) {
                ZKW_CUDA(ctx, cudaEventRecord(ctx->lane_done[i], ctx->lane_stream[i]));
                ZKW_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->lane_done[i], 0));
            }
        }
        ZKW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < m; i++) msm_collect(ctx, i, outs[done + i]);
        done += m;
    }
    return ZKW_OK;
}

int msm_lane_submit(zkw_ctx* ctx, int lane, const MsmJob& job) {
    if (lane < 1 || lane >= zkw_ctx::kMsmLanes) return ZKW_ERR_INVALID;
    ZKW_TRY(lane_init(ctx, 0));
    ZKW_TRY(lane_init(ctx, lane));
    ZKW_CUDA(ctx, cudaEventRecord(ctx->fork_event, ctx->stream));
    ZKW_CUDA(ctx, cudaStreamWaitEvent(ctx->lane_stream[lane], ctx->fork_event, 0));
    ZKW_TRY(msm_enqueue(ctx, lane, job.which_bases, job.bases_dev, job.scalars_dev, job.n));
    ZKW_CUDA(ctx, cudaEventRecord(ctx->lane_done[lane], ctx->lane_stream[lane]));
    return ZKW_OK;
}

int msm_lane_wait(zkw_ctx* ctx, int lane, uint64_t out_xyz[12]) {
    if (lane < 1 || lane >= zkw_ctx::kMsmLanes || !ctx->lane_done[lane]) return ZKW_ERR_INVALID;
    ZKW_CUDA(ctx, cudaEventSynchronize(ctx->lane_done[lane]));
    msm_collect(ctx, lane, out_xyz);
    return ZKW_OK;
}

int msm_lanes_collect(zkw_ctx* ctx, const int* lanes, int count, uint64_t (*outs)[12]) {
    for (int i = 0; i < count; i++) ZKW_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->lane_done[lanes[i]], 0));
    ZKW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < count; i++) msm_collect(ctx, lanes[i], outs[i]);
    return ZKW_OK;
}

}  // namespace zkw
