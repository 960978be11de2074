// prover_kernels.cuh — device kernels for the parts of halo2's create_proof that sit between the MSM /
// NTT / quotient calls: blinding rows, the lookup argument's permuted columns, the
// permutation / lookup grand products, polynomial evaluations at the challenge points, the GWC batching
// and the (X - z) divisions.  All of it is streaming integer work over n x 32-byte vectors.
//
// Upstream functions these replace (halo2_proofs, PSE v2023_01_20; reached from the reference through
// create_proof, halo2-circuits/src/ecc/ecdsa_p256.rs:366-373, 416-423, 555-562):
//   plonk::lookup::prover::{commit_permuted -> permute_expression_pair, commit_product}
//   plonk::permutation::prover::Argument::commit
//   arithmetic::{eval_polynomial, kate_division}, poly::kzg::multiopen::gwc::ProverGWC::create_proof
#pragma once
#include "common.cuh"

namespace zkw {

constexpr int kScanThreads = 256;
constexpr int kScanPerThread = 8;
constexpr int kScanBlock = kScanThreads * kScanPerThread;  // 2048 elements per CTA

// ---- blinding stream (mirrors oracle/halo2_ref.py::rand_fr) -------------------------------------------------
// Upstream draws every blinding scalar and the vanishing argument's random polynomial from OsRng
// (ecdsa_p256.rs:362,412).  Here: ChaCha20 keyed with a 256-bit seed (drawn from the OS by the caller), nonce =
// the stream id (one stream per blinded column), block counter = the element index.  One 64-byte block gives 512
// uniform bits, reduced modulo r (bias 2^-258).  A fixed key reproduces a proof bit for bit - what the parity tests
// against the oracle prover use.
struct RandKey { uint32_t w[8]; };

__host__ __device__ inline uint32_t rotl32(uint32_t x, int c) { return (x << c) | (x >> (32 - c)); }

__host__ __device__ inline void chacha20_block(const RandKey& key, uint64_t stream, uint64_t counter, uint32_t out[16]) {
    uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key.w[0], key.w[1], key.w[2], key.w[3],
                      key.w[4], key.w[5], key.w[6], key.w[7], (uint32_t)counter, (uint32_t)(counter >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
    uint32_t x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = s[i];
#define ZKW_QR(a, b, c, d)                                                                                             \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 12);                        \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 8);  x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 7);
#pragma unroll 1
    for (int r = 0; r < 10; r++) {
        ZKW_QR(0, 4, 8, 12) ZKW_QR(1, 5, 9, 13) ZKW_QR(2, 6, 10, 14) ZKW_QR(3, 7, 11, 15)
        ZKW_QR(0, 5, 10, 15) ZKW_QR(1, 6, 11, 12) ZKW_QR(2, 7, 8, 13) ZKW_QR(3, 4, 9, 14)
    }
#undef ZKW_QR
#pragma unroll
    for (int i = 0; i < 16; i++) out[i] = x[i] + s[i];
}

// (lo + 2^256 hi) mod r in Montgomery form, lo = words 0..7, hi = words 8..15 (little-endian)
__host__ __device__ inline Fr rand_fr(const RandKey& key, uint64_t stream, uint64_t index) {
    uint32_t b[16];
    chacha20_block(key, stream, index, b);
    Fr lo, hi;
#pragma unroll
    for (int j = 0; j < 8; j++) { lo.l[j] = b[j]; hi.l[j] = b[8 + j]; }
    for (int j = 0; j < 5; j++) { lo.reduce_once(); hi.reduce_once(); }   // 2^256 < 6r
    const Fr r2 = Fr::r2();
    return lo * r2 + hi * (r2 * r2);      // lo R + hi R^2 (mod r) = Montgomery form of lo + hi 2^256
}

__global__ void rand_fill_kernel(uint4* out, size_t count, RandKey key, uint64_t stream, uint64_t first_index) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    rand_fr(key, stream, first_index + i).store(out + 2 * i);
}

__global__ void zero_fill_kernel(uint4* out, size_t count) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    out[2 * i] = make_uint4(0, 0, 0, 0);
    out[2 * i + 1] = make_uint4(0, 0, 0, 0);
}

// canonical little-endian integers (< 2r) -> Montgomery form, in place
__global__ void to_mont_kernel(uint4* v, size_t count) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    Fr x = Fr::load(v + 2 * i);
    x.reduce_once();
    x.to_mont().store(v + 2 * i);
}

// one u64 per row -> Montgomery field elements (in and out must not overlap)
__global__ void u64_to_mont_kernel(const uint64_t* in, uint4* out, size_t count) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint64_t v = in[i];
    Fr x = Fr::zero();
    x.l[0] = (uint32_t)v;
    x.l[1] = (uint32_t)(v >> 32);
    x.to_mont().store(out + 2 * i);
}

// out[i] = a[i] * b[i]
__global__ void mul_vec_kernel(const uint4* a, const uint4* b, uint4* out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    (Fr::load(a + 2 * i) * Fr::load(b + 2 * i)).store(out + 2 * i);
}

// omega^i from the half-size twiddle table (omega^(i + n/2) = -omega^i)
__device__ __forceinline__ Fr omega_pow(const uint4* tw, size_t i, size_t half) {
    Fr w = Fr::load_nc(tw + 2 * (i & (half - 1)));
    return (i & half) ? w.neg() : w;
}

// ---- grand-product numerators / denominators ------------------------------------------------------------
struct PermChunkArgs {
    const uint4* values[8];   // column values (Lagrange), this chunk
    const uint4* sigmas[8];   // sigma values (Lagrange), this chunk
    Fr delta_beta[8];         // delta^(global column index) * beta
    int ncols;
    Fr beta, gamma;
    const uint4* tw;          // omega^i, i < n/2
    size_t n;
};

// num[i] = prod_c (v_c + delta^c beta omega^i + gamma), den[i] = prod_c (v_c + beta sigma_c + gamma)
__global__ void __launch_bounds__(128) perm_numden_kernel(const PermChunkArgs a, uint4* num, uint4* den) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const Fr w = omega_pow(a.tw, i, a.n >> 1);
    Fr nu = Fr::one(), de = Fr::one();
    for (int c = 0; c < a.ncols; c++) {
        const Fr v = Fr::load(a.values[c] + 2 * i);
        const Fr s = Fr::load_nc(a.sigmas[c] + 2 * i);
        nu = nu * (v + a.delta_beta[c] * w + a.gamma);
        de = de * (v + a.beta * s + a.gamma);
    }
    nu.store(num + 2 * i);
    de.store(den + 2 * i);
}

// lookup: num = (input + beta)(table + gamma), den = (A' + beta)(S' + gamma)
__global__ void __launch_bounds__(128) lookup_numden_kernel(const uint4* inp, const uint4* tab, const uint4* ap, const uint4* sp,
                                                            Fr beta, Fr gamma, uint4* num, uint4* den, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ((Fr::load(inp + 2 * i) + beta) * (Fr::load_nc(tab + 2 * i) + gamma)).store(num + 2 * i);
    ((Fr::load(ap + 2 * i) + beta) * (Fr::load(sp + 2 * i) + gamma)).store(den + 2 * i);
}

// ---- three-phase scans over field elements ------------------------------------------------------------------
// MUL = true: running products, false: running sums.  REVERSE = true scans from the last element down.
// Inclusive: out[i] = x[0] op ... op x[i]  (or x[i] op ... op x[n-1] when REVERSE).
template <bool MUL>
__device__ __forceinline__ Fr scan_op(const Fr& a, const Fr& b) { return MUL ? a * b : a + b; }
template <bool MUL>
__device__ __forceinline__ Fr scan_identity() { return MUL ? Fr::one() : Fr::zero(); }

// phase 1: per-CTA totals
template <bool MUL, bool REVERSE>
__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const uint4* x, uint4* block_totals, size_t n) {
    __shared__ uint4 sh[kScanThreads * 2];
    const size_t base = (size_t)blockIdx.x * kScanBlock + (size_t)threadIdx.x * kScanPerThread;
    Fr acc = scan_identity<MUL>();
#pragma unroll
    for (int j = 0; j < kScanPerThread; j++) {
        size_t i = base + j;
        if (i < n) {
            size_t idx = REVERSE ? n - 1 - i : i;
            acc = scan_op<MUL>(acc, Fr::load(x + 2 * idx));
        }
    }
    acc.store(sh + 2 * threadIdx.x);
    __syncthreads();
    for (int d = kScanThreads / 2; d >= 1; d >>= 1) {
        if (threadIdx.x < d) {
            Fr a = Fr::load(sh + 2 * threadIdx.x), b = Fr::load(sh + 2 * (threadIdx.x + d));
            scan_op<MUL>(a, b).store(sh + 2 * threadIdx.x);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        block_totals[2 * blockIdx.x] = sh[0];
        block_totals[2 * blockIdx.x + 1] = sh[1];
    }
}

// phase 2: one CTA turns the block totals into exclusive block prefixes (in place); also emits the grand total
template <bool MUL>
__global__ void __launch_bounds__(1024) scan_blocks_kernel(uint4* block_totals, size_t nblocks, uint4* grand_total) {
    __shared__ uint4 sh[1024 * 2];
    const int t = threadIdx.x;
    const size_t per = (nblocks + 1023) / 1024;
    const size_t lo = (size_t)t * per, hi = lo + per < nblocks ? lo + per : nblocks;
    Fr acc = scan_identity<MUL>();
    for (size_t i = lo; i < hi; i++) acc = scan_op<MUL>(acc, Fr::load(block_totals + 2 * i));
    acc.store(sh + 2 * t);
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        Fr v = scan_identity<MUL>();
        const bool has = t >= d;
        if (has) v = Fr::load(sh + 2 * (t - d));
        __syncthreads();
        if (has) scan_op<MUL>(v, Fr::load(sh + 2 * t)).store(sh + 2 * t);
        __syncthreads();
    }
    Fr run = t == 0 ? scan_identity<MUL>() : Fr::load(sh + 2 * (t - 1));  // exclusive prefix of this thread's run
    for (size_t i = lo; i < hi; i++) {
        Fr cur = Fr::load(block_totals + 2 * i);
        run.store(block_totals + 2 * i);
        run = scan_op<MUL>(run, cur);
    }
    if (t == 1023 && grand_total) {
        grand_total[0] = sh[2 * 1023];
        grand_total[1] = sh[2 * 1023 + 1];
    }
}

// phase 3: inclusive scan within the CTA, offset by the CTA's exclusive prefix
template <bool MUL, bool REVERSE>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const uint4* x, const uint4* block_prefix, uint4* out, size_t n) {
    __shared__ uint4 sh[kScanThreads * 2];
    const size_t base = (size_t)blockIdx.x * kScanBlock + (size_t)threadIdx.x * kScanPerThread;
    Fr v[kScanPerThread];
    Fr acc = scan_identity<MUL>();
#pragma unroll
    for (int j = 0; j < kScanPerThread; j++) {
        size_t i = base + j;
        if (i < n) {
            size_t idx = REVERSE ? n - 1 - i : i;
            acc = scan_op<MUL>(acc, Fr::load(x + 2 * idx));
        }
        v[j] = acc;
    }
    acc.store(sh + 2 * threadIdx.x);
    __syncthreads();
    for (int d = 1; d < kScanThreads; d <<= 1) {
        Fr u = scan_identity<MUL>();
        const bool has = (int)threadIdx.x >= d;
        if (has) u = Fr::load(sh + 2 * (threadIdx.x - d));
        __syncthreads();
        if (has) scan_op<MUL>(u, Fr::load(sh + 2 * threadIdx.x)).store(sh + 2 * threadIdx.x);
        __syncthreads();
    }
    Fr pre = Fr::load(block_prefix + 2 * blockIdx.x);
    if (threadIdx.x > 0) pre = scan_op<MUL>(pre, Fr::load(sh + 2 * (threadIdx.x - 1)));
#pragma unroll
    for (int j = 0; j < kScanPerThread; j++) {
        size_t i = base + j;
        if (i < n) {
            size_t idx = REVERSE ? n - 1 - i : i;
            scan_op<MUL>(pre, v[j]).store(out + 2 * idx);
        }
    }
}

// z[0] = z0; z[i+1] = z0 * PN[i] * SD[i+1] * Tinv for i+1 <= u (SD[n] = 1); rows above u are blinding rows
// PN = inclusive prefix products of num, SD = inclusive suffix products of den, T = SD[0].
__global__ void __launch_bounds__(128) grand_product_finalize_kernel(const uint4* pn, const uint4* sd, const uint4* z0_ptr, Fr tinv,
                                                                     uint4* z, size_t u, size_t n) {
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > u) return;
    const Fr z0 = z0_ptr ? Fr::load(z0_ptr) : Fr::one();
    if (r == 0) { z0.store(z); return; }
    Fr s = r < n ? Fr::load(sd + 2 * r) : Fr::one();
    (z0 * tinv * Fr::load(pn + 2 * (r - 1)) * s).store(z + 2 * r);
}

// ---- lookup argument: permuted input / table ---------------------------------------------------------------
// 256-bit compare of canonical little-endian values
__device__ __forceinline__ int cmp256(const Fr& a, const Fr& b) {
#pragma unroll
    for (int i = 7; i >= 0; i--) {
        if (a.l[i] != b.l[i]) return a.l[i] < b.l[i] ? -1 : 1;
    }
    return 0;
}

// rank[i] = index of input value i in the sorted distinct table (canonical form); histogram of ranks
__global__ void lookup_rank_kernel(const uint4* inp, const uint4* table_sorted_canon, uint32_t m, uint32_t* rank, uint32_t* counts,
                                   uint32_t* error_flag, size_t u, int probe) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    // every lane stays to the end: the histogram add is aggregated per warp (below)
    uint32_t found = 0xffffffffu;   // rank of this row's value, or none
    if (i < u) {
        const Fr v = Fr::load(inp + 2 * i).from_mont();
        // range tables hold 0 .. T-1 in order, so a value below m is (almost always) its own rank: one probe instead of log2(m)
        // dependent loads; anything else takes the binary search
        if (probe && !(v.l[1] | v.l[2] | v.l[3] | v.l[4] | v.l[5] | v.l[6] | v.l[7]) && v.l[0] < m &&
            cmp256(Fr::load_nc(table_sorted_canon + 2 * (size_t)v.l[0]), v) == 0) {
            found = v.l[0];
        } else {
            uint32_t lo = 0, hi = m;  // first index with table[idx] >= v
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (cmp256(Fr::load_nc(table_sorted_canon + 2 * (size_t)mid), v) < 0) lo = mid + 1; else hi = mid;
            }
            if (lo >= m || cmp256(Fr::load_nc(table_sorted_canon + 2 * (size_t)lo), v) != 0) {
                atomicExch(error_flag, 1u);  // ConstraintSystemFailure upstream: lookup input not in table
                rank[i] = 0;
            } else {
                found = lo;
            }
        }
        if (found != 0xffffffffu) rank[i] = found;
    }
    // The lookup input of the ECDSA circuit is q_lookup * advice: seven rows of eight hold 0, so one address took 460 000 of the
    // 2^19 atomic adds and L2 serialised them (0.32 ms for a kernel whose loads take 10 us).  Lanes that agree on the rank
    // find each other (match.any) and the lowest one adds for the group.
    __syncwarp();
    const unsigned have = __ballot_sync(0xffffffffu, found != 0xffffffffu);
    if (found != 0xffffffffu) {
        const unsigned peers = __match_any_sync(have, found);
        if ((threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&counts[found], (uint32_t)__popc(peers));
    }
}

// three exclusive scans over the m distinct table values in one single-CTA kernel:
//   run_start[j]  = sum_{j'<j} c[j']                       (start row of value j's run in A')
//   rep_start[j]  = sum_{j'<j} max(c[j']-1, 0)              (index of its first repeated row)
//   desc_start[q] = sum_{q'<q} left[m-1-q'],  left[j] = mult[j] - (c[j] > 0)   (leftovers, descending value order)
// Tiles of 1024 entries, warp-shuffle scans, running carry (coalesced loads).
__device__ __forceinline__ uint32_t warp_incl_scan_u32(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

// Two launches over tiles of 4096 entries (four consecutive entries per thread, one CTA per tile): the tile totals, then every
// CTA sums the totals of the tiles before it and scans its own tile.  (One CTA walking all 64 tiles of a 2^18-entry table with
// barriers in between took 0.32 ms on the proof's critical path between the lookup ranks and A' / S'.)
constexpr uint32_t kLookupTile = 4096;
__device__ __forceinline__ void lookup_scan_entry(const uint32_t* counts, const uint32_t* mult, uint32_t m, uint32_t j, uint32_t q[3],
                                                  uint32_t* error_flag) {
    const uint32_t cj = counts[j];
    q[0] = cj;
    q[1] = cj ? cj - 1 : 0;
    const uint32_t jj = m - 1 - j;
    const uint32_t cjj = counts[jj], mu = mult[jj];
    if (cjj && mu == 0) atomicExch(error_flag, 1u);
    q[2] = mu - (cjj ? 1u : 0u);
}

__global__ void __launch_bounds__(1024) lookup_scan_totals_kernel(const uint32_t* counts, const uint32_t* mult, uint32_t m,
                                                                  uint32_t* tile_tot /* [3][tiles] */, uint32_t* error_flag) {
    __shared__ uint32_t wsum[3][32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    uint32_t tot[3] = {0, 0, 0};
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const uint32_t j = blockIdx.x * kLookupTile + 4 * (uint32_t)t + e;
        if (j < m) {
            uint32_t q[3];
            lookup_scan_entry(counts, mult, m, j, q, error_flag);
            tot[0] += q[0]; tot[1] += q[1]; tot[2] += q[2];
        }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int d = 16; d; d >>= 1) tot[k] += __shfl_xor_sync(0xffffffffu, tot[k], d);
        if (lane == 0) wsum[k][warp] = tot[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            uint32_t v = wsum[k][lane];
#pragma unroll
            for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
            if (lane == 0) tile_tot[(size_t)k * gridDim.x + blockIdx.x] = v;
        }
    }
}

__global__ void __launch_bounds__(1024) lookup_scan_kernel(const uint32_t* counts, const uint32_t* mult, uint32_t m, const uint32_t* tile_tot,
                                                           uint32_t* run_start, uint32_t* rep_start, uint32_t* desc_start, uint32_t* error_flag) {
    __shared__ uint32_t wsum[3][32];
    __shared__ uint32_t carry[3];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t tiles = gridDim.x;
    if (warp < 3) {   // warp k: total of the tiles before this one, array k
        uint32_t v = 0;
        for (uint32_t b = lane; b < blockIdx.x; b += 32) v += tile_tot[(size_t)warp * tiles + b];
#pragma unroll
        for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if (lane == 0) carry[warp] = v;
    }
    uint32_t q[3][4];
    uint32_t tot[3] = {0, 0, 0};
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const uint32_t j = blockIdx.x * kLookupTile + 4 * (uint32_t)t + e;
        q[0][e] = q[1][e] = q[2][e] = 0;
        if (j < m) {
            uint32_t qq[3];
            lookup_scan_entry(counts, mult, m, j, qq, error_flag);
            q[0][e] = qq[0]; q[1][e] = qq[1]; q[2][e] = qq[2];
        }
#pragma unroll
        for (int k = 0; k < 3; k++) tot[k] += q[k][e];
    }
    uint32_t inc[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        inc[k] = warp_incl_scan_u32(tot[k], lane);
        if (lane == 31) wsum[k][warp] = inc[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) wsum[k][lane] = warp_incl_scan_u32(wsum[k][lane], lane);
    }
    __syncthreads();
    uint32_t* outs[3] = {run_start, rep_start, desc_start};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        uint32_t run = carry[k] + (warp ? wsum[k][warp - 1] : 0u) + inc[k] - tot[k];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const uint32_t j = blockIdx.x * kLookupTile + 4 * (uint32_t)t + e;
            if (j < m) { outs[k][j] = run; run += q[k][e]; }
        }
        if (blockIdx.x == tiles - 1 && t == 1023) outs[k][m] = carry[k] + wsum[k][31];   // grand totals
    }
}

// last index q in [0, m) with start[q] <= x  (starts non-decreasing, start[m] = total > x)
__device__ __forceinline__ uint32_t last_le(const uint32_t* start, uint32_t m, uint32_t x) {
    uint32_t lo = 0, hi = m;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (start[mid] <= x) lo = mid; else hi = mid;
    }
    return lo;
}

// A'[r] = value of the run containing row r; S'[r] = that value at the run's first row, else the
// (rep index)-th leftover table value in DESCENDING value order (upstream pops repeated rows from the end
// while walking the leftovers in ascending order).
__global__ void lookup_expand_kernel(const uint4* table_sorted_mont, uint32_t m, const uint32_t* run_start, const uint32_t* rep_start,
                                     const uint32_t* desc_start, uint4* a_out, uint4* s_out, size_t u) {
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= u) return;
    const uint32_t j = last_le(run_start, m, (uint32_t)r);
    const uint4 v0 = table_sorted_mont[2 * (size_t)j], v1 = table_sorted_mont[2 * (size_t)j + 1];
    a_out[2 * r] = v0; a_out[2 * r + 1] = v1;
    const uint32_t off = (uint32_t)r - run_start[j];
    if (off == 0) {
        s_out[2 * r] = v0; s_out[2 * r + 1] = v1;
    } else {
        const uint32_t q = rep_start[j] + off - 1;
        const uint32_t dq = last_le(desc_start, m, q);
        const uint32_t jj = m - 1 - dq;
        s_out[2 * r] = table_sorted_mont[2 * (size_t)jj];
        s_out[2 * r + 1] = table_sorted_mont[2 * (size_t)jj + 1];
    }
}

// ---- batched polynomial evaluation -------------------------------------------------------------------------
// One (poly, point) pair per blockIdx.y.  Level 0: each CTA reduces kScanBlock coefficients to
// sum_i c_i x^(i - chunk_start) using per-thread Horner and a tree over threads with x^(8*2^l); the
// per-CTA partials form a polynomial in x^2048 that the same kernel reduces at the next level.
struct EvalJob {
    const uint4* coeffs;   // level-0 input
    uint32_t point;        // index into the power tables
};
// pow_table[point][l] = x^(2^l), l < 40
__global__ void __launch_bounds__(kScanThreads) eval_reduce_kernel(const EvalJob* jobs, const uint4* level_base, size_t in_stride,
                                                                  const uint4* pow_table, int log_stride, uint4* partial_out, size_t n,
                                                                  size_t out_stride) {
    __shared__ uint4 sh[kScanThreads * 2];
    const EvalJob job = jobs[blockIdx.y];
    const uint4* src = level_base ? level_base + 2 * (size_t)blockIdx.y * in_stride : job.coeffs;
    const uint4* pw = pow_table + 2 * 40 * (size_t)job.point;
    const Fr x1 = Fr::load(pw + 2 * log_stride);  // x^(2^log_stride): the variable at this level
    const size_t base = (size_t)blockIdx.x * kScanBlock + (size_t)threadIdx.x * kScanPerThread;
    // Horner on purpose: the eight-term Fr::dot_lazy form (576 wide multiplies, no dependency chain, against 1024) was measured
    // SLOWER here - 0.340 ms against 0.262 ms for the 18 evaluations of a k = 19 proof - because its 118 registers leave two CTAs
    // per SM where this 40-register loop runs eight, and the kernel lives on thread-level parallelism, not on the multiplier
    Fr acc = Fr::zero();
#pragma unroll
    for (int j = kScanPerThread - 1; j >= 0; j--) {
        size_t i = base + j;
        Fr c = i < n ? Fr::load(src + 2 * i) : Fr::zero();
        acc = acc * x1 + c;
    }
    acc.store(sh + 2 * threadIdx.x);
    __syncthreads();
    // tree: pair (t, t + d) with weight x1^(8 d); 8 = 2^3 coefficients per thread
    int lvl = log_stride + 3;
    for (int d = 1; d < kScanThreads; d <<= 1, lvl++) {
        if ((threadIdx.x & (2 * d - 1)) == 0) {
            Fr a = Fr::load(sh + 2 * threadIdx.x), b = Fr::load(sh + 2 * (threadIdx.x + d));
            (a + b * Fr::load(pw + 2 * lvl)).store(sh + 2 * threadIdx.x);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        uint4* o = partial_out + 2 * ((size_t)blockIdx.y * out_stride + blockIdx.x);
        o[0] = sh[0];
        o[1] = sh[1];
    }
}

// ---- linear combination: out[i] = sum_j w_j * p_j[i] ---------------------------------------------------------
struct LinCombArgs {
    const uint4* const* polys;  // device array of npolys pointers
    const uint4* weights;       // device array of npolys field elements
    int npolys;
    size_t n;
};
__global__ void __launch_bounds__(128) lincomb_kernel(const LinCombArgs a, uint4* out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    Fr acc = Fr::zero();
    int j = 0;
    for (; j + 4 <= a.npolys; j += 4) {   // four terms per Montgomery pass (Fr::dot_lazy: one reduction for the four products)
        Fr w[4], v[4];
#pragma unroll
        for (int t = 0; t < 4; t++) { w[t] = Fr::load_nc(a.weights + 2 * (j + t)); v[t] = Fr::load(a.polys[j + t] + 2 * i); }
        acc = acc + Fr::dot_lazy<4>(w, v).normalized();
    }
    for (; j < a.npolys; j++) acc = acc + Fr::load_nc(a.weights + 2 * j) * Fr::load(a.polys[j] + 2 * i);
    acc.store(out + 2 * i);
}

// poly[d] -= low[d] for d < m (m <= 8): subtracting a low-degree polynomial
__global__ void sub_low_kernel(uint4* poly, const uint4* low, int m) {
    const int d = threadIdx.x;
    if (d >= m) return;
    (Fr::load(poly + 2 * d) - Fr::load(low + 2 * d)).store(poly + 2 * d);
}

// ---- (p(X) - p(z)) / (X - z) --------------------------------------------------------------------------------
// q_{i-1} = sum_{j >= i} a_j z^(j-i) = z^-i (p(z) - sum_{j<i} a_j z^j).  Step 1: t_j = a_j z^j.
// Step 2: additive prefix scan of t (scan_*<false,false>).  Step 3: q_{i-1} = z^-i (E - PRE_{i-1}),
// E = PRE_{n-1} = p(z).  z^j and z^-j are rebuilt per thread from a power seed (pow of a 64-bit exponent).
__global__ void __launch_bounds__(128) kate_terms_kernel(const uint4* a, uint4* t, Fr z, size_t n) {
    const size_t start = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (start >= n) return;
    Fr p = z.pow((uint64_t)start);
    for (size_t i = start; i < start + 16 && i < n; i++) {
        (Fr::load(a + 2 * i) * p).store(t + 2 * i);
        p = p * z;
    }
}
// pre = inclusive prefix sums of t; q has n-1 entries: q[i-1] = zinv^i * (pre[n-1] - pre[i-1]) for i = 1..n-1
__global__ void __launch_bounds__(128) kate_finish_kernel(const uint4* pre, uint4* q, Fr zinv, size_t n) {
    const size_t start = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16 + 1;
    if (start >= n) return;
    const Fr e = Fr::load(pre + 2 * (n - 1));
    Fr p = zinv.pow((uint64_t)start);
    for (size_t i = start; i < start + 16 && i < n; i++) {
        ((e - Fr::load(pre + 2 * (i - 1))) * p).store(q + 2 * (i - 1));
        p = p * zinv;
    }
}

// sigma values from the permutation mapping: sigma[c][r] = delta^(c') * omega^(r'), mapping as (c', r') u32 pairs
__global__ void sigma_values_kernel(const uint2* mapping, const uint4* delta_pows, const uint4* tw, uint4* out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint2 m = mapping[i];
    (Fr::load_nc(delta_pows + 2 * (size_t)m.x) * omega_pow(tw, m.y, n >> 1)).store(out + 2 * i);
}

}  // namespace zkw
