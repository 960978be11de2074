// prover.cu — halo2's keygen + create_proof for the P-256 ECDSA circuit's constraint system, run on the
// device end to end: the replacement of what the reference reaches through
//   download_keys   -> keygen_vk / keygen_pk              halo2-circuits/src/ecc/ecdsa_p256.rs:256-272
//   generate_proof_evm -> create_proof<.., ProverGWC, .., EvmTranscript, ..>      ecdsa_p256.rs:329-377
//   generate_proof     -> create_proof<.., Blake2bWrite, ..>                      ecdsa_p256.rs:379-427
// (PSE halo2_proofs v2023_01_20 semantics, restated independently in oracle/halo2_ref.py, whose verifier
// accepts the reference's golden proof).  The host code here only sequences kernels and runs the
// transcript; every vector stays in HBM from the H2D copy of the witness to the D2H copy of the
// commitments / evaluations.  Multi-open: GWC (generate_proof_evm) or SHPLONK (generate_proof, ZKW_MULTIOPEN_SHPLONK),
// under either transcript.
#include <algorithm>
#include <array>
#include <cstring>
#include <memory>
#include "common.cuh"
#include "hash.hpp"
#include "prover_kernels.cuh"

namespace zkw {

static const uint64_t kDeltaM[4] = {0x9a0c322befd78855ULL, 0x46e82d14249b563cULL, 0x5983a663e0b0b7a7ULL, 0x22ab452baaa111adULL};

static Fr fr_of(const uint64_t v[4]) { Fr r; memcpy(r.l, v, 32); return r; }

// ---- transcripts -----------------------------------------------------------------------------------
struct Transcript {
    int kind;  // 0 = Blake2b / Challenge255, 1 = EVM (keccak)
    Blake2b b2{64, "Halo2-Transcript"};
    std::vector<uint8_t> buf;   // EVM running buffer
    std::vector<uint8_t> out;   // proof bytes

    explicit Transcript(int k) : kind(k) {}

    static void canon_le(const Fr& m, uint8_t o[32]) { Fr c = m.from_mont(); memcpy(o, c.l, 32); }
    static void canon_le_q(const Fq& m, uint8_t o[32]) { Fq c = m.from_mont(); memcpy(o, c.l, 32); }
    static void reverse32(uint8_t o[32]) { std::reverse(o, o + 32); }

    void common_scalar(const Fr& s) {
        uint8_t b[32];
        canon_le(s, b);
        if (kind == 0) { b2.update_byte(2); b2.update(b, 32); }
        else { reverse32(b); buf.insert(buf.end(), b, b + 32); }
    }
    void write_scalar(const Fr& s) {
        common_scalar(s);
        uint8_t b[32];
        canon_le(s, b);
        if (kind == 1) reverse32(b);
        out.insert(out.end(), b, b + 32);
    }
    // pt: affine Montgomery (x, y); the identity never occurs for commitments to non-zero polynomials
    void write_point(const uint64_t xy[8]) {
        Fq x, y;
        memcpy(x.l, xy, 32);
        memcpy(y.l, xy + 4, 32);
        uint8_t bx[32], by[32];
        canon_le_q(x, bx);
        canon_le_q(y, by);
        if (kind == 0) {
            b2.update_byte(1);
            b2.update(bx, 32);
            b2.update(by, 32);
            uint8_t c[32];
            memcpy(c, bx, 32);
            c[31] |= (uint8_t)((by[0] & 1) << 7);   // halo2curves: x little-endian, bit 255 = parity of y
            out.insert(out.end(), c, c + 32);
        } else {
            reverse32(bx);
            reverse32(by);
            buf.insert(buf.end(), bx, bx + 32);
            buf.insert(buf.end(), by, by + 32);
            out.insert(out.end(), bx, bx + 32);
            out.insert(out.end(), by, by + 32);
        }
    }
    Fr squeeze() {
        if (kind == 0) {
            b2.update_byte(0);
            uint8_t h[64];
            b2.peek_digest(h);
            Fr lo, hi;
            memcpy(lo.l, h, 32);
            memcpy(hi.l, h + 32, 32);
            const Fr r2 = Fr::r2();
            return lo * r2 + (hi * r2) * r2;   // (lo + hi 2^256) mod r, in Montgomery form
        }
        std::vector<uint8_t> data = buf;
        if (buf.size() == 32) data.push_back(1);
        uint8_t h[32];
        keccak256(data.data(), data.size(), h);
        buf.assign(h, h + 32);
        uint8_t le[32];
        memcpy(le, h, 32);
        reverse32(le);
        Fr v;
        memcpy(v.l, le, 32);
        return v * Fr::r2();
    }
};

}  // namespace zkw

using namespace zkw;

// ---- proving key ------------------------------------------------------------------------------------
struct zkw_pk {
    zkw_circuit_shape shape;
    size_t n = 0, en = 0, u = 0;
    unsigned A = 0, L = 0, F = 0, nfixed = 0, nperm = 0, nsets = 0, nlk = 0, chunk = 0;
    std::vector<uint64_t*> fixed_values, fixed_polys, fixed_cosets;
    std::vector<uint64_t*> sigma_values, sigma_polys, sigma_cosets;
    uint64_t *l0_coset = nullptr, *l_last_coset = nullptr, *l_active_coset = nullptr;
    uint64_t *table_canon = nullptr, *table_mont = nullptr;
    uint32_t* table_mult = nullptr;
    uint32_t table_m = 0;
    std::vector<std::array<uint64_t, 8>> fixed_commitments, perm_commitments;
    uint64_t digest[4] = {0, 0, 0, 0};
    std::vector<void*> owned;
    unsigned table_col() const { return F; }
    unsigned q_enable_col(unsigned c) const { return F + 1 + c; }
    unsigned q_lookup_col() const { return F + 1 + A; }
};

namespace zkw {

static int dmalloc(zkw_ctx* ctx, zkw_pk* pk, size_t bytes, void** out) {
    ZKW_CUDA(ctx, cudaMalloc(out, bytes ? bytes : 1));
    if (pk) pk->owned.push_back(*out);
    return ZKW_OK;
}

// Per-proof device memory: bump allocation out of one arena owned by the context (stack discipline for
// nested scopes).  The first proof of a given shape falls back to cudaMalloc for what does not fit and
// the arena is grown to the high-water mark afterwards, so steady-state proofs allocate nothing.
struct Scratch {
    zkw_ctx* ctx;
    size_t mark_off, mark_virtual;
    std::vector<void*> extra;
    explicit Scratch(zkw_ctx* c) : ctx(c), mark_off(c->arena_off), mark_virtual(c->arena_virtual) {}
    ~Scratch() {
        cudaStreamSynchronize(ctx->stream);
        for (int i = 0; i < zkw_ctx::kAuxStreams; i++) if (ctx->aux_stream[i]) cudaStreamSynchronize(ctx->aux_stream[i]);
        for (int i = 1; i < zkw_ctx::kMsmLanes; i++) if (ctx->lane_stream[i]) cudaStreamSynchronize(ctx->lane_stream[i]);
        for (void* p : extra) cudaFree(p);
        ctx->arena_off = mark_off;
        ctx->arena_virtual = mark_virtual;
        if (mark_virtual == 0 && ctx->arena_need > ctx->arena.bytes) {
            if (ctx->arena.ptr) cudaFree(ctx->arena.ptr);
            ctx->arena.ptr = nullptr;
            ctx->arena.bytes = 0;
            if (cudaMalloc(&ctx->arena.ptr, ctx->arena_need) == cudaSuccess) ctx->arena.bytes = ctx->arena_need;
            else { ctx->arena.ptr = nullptr; cudaGetLastError(); }
        }
    }
    int get(size_t bytes, void** out) {
        bytes = (bytes + 255) & ~(size_t)255;
        if (!bytes) bytes = 256;
        ctx->arena_virtual += bytes;
        if (ctx->arena_virtual > ctx->arena_need) ctx->arena_need = ctx->arena_virtual;
        if (ctx->arena.ptr && ctx->arena_off + bytes <= ctx->arena.bytes) {
            *out = (char*)ctx->arena.ptr + ctx->arena_off;
            ctx->arena_off += bytes;
            return ZKW_OK;
        }
        ZKW_CUDA(ctx, cudaMalloc(out, bytes));
        extra.push_back(*out);
        return ZKW_OK;
    }
};

static unsigned grid_for(size_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }

static int commit_dev(zkw_ctx* ctx, int which, const uint64_t* scalars_dev, size_t n, uint64_t out_xy[8]) {
    uint64_t xyz[12];
    ZKW_TRY(msm_run(ctx, which, nullptr, scalars_dev, n, xyz));
    memcpy(out_xy, xyz, 64);
    bool ident = true;
    for (int i = 8; i < 12; i++) ident = ident && xyz[i] == 0;
    if (ident) memset(out_xy, 0, 64);
    return ZKW_OK;
}

// commit several polynomials at once (independent MSMs overlap on the context's lanes) and append the
// commitments to the transcript in order
static int commit_batch(zkw_ctx* ctx, Transcript& tr, const std::vector<std::pair<int, const uint64_t*>>& polys, size_t n) {
    std::vector<MsmJob> jobs(polys.size());
    for (size_t i = 0; i < polys.size(); i++) jobs[i] = MsmJob{polys[i].first, nullptr, polys[i].second, n};
    std::vector<std::array<uint64_t, 12>> outs(polys.size());
    ZKW_TRY(msm_run_batch(ctx, jobs.data(), (int)jobs.size(), reinterpret_cast<uint64_t(*)[12]>(outs.data())));
    for (auto& o : outs) {
        uint64_t xy[8];
        memcpy(xy, o.data(), 64);
        bool ident = true;
        for (int i = 8; i < 12; i++) ident = ident && o[i] == 0;
        if (ident) memset(xy, 0, 64);
        tr.write_point(xy);
    }
    return ZKW_OK;
}

// Pipelined commitments: each polynomial's MSM is submitted to a side lane the moment the polynomial is
// queued on the main stream, which keeps producing the next one.  submit() returns a ticket; write(ticket)
// waits for that MSM only and appends its commitment to the transcript, so commitments whose polynomials do
// not depend on a challenge (permuted lookup columns, the random polynomial) can be computed rounds before
// the transcript absorbs them.  When every lane is busy the oldest ticket is retired into `done` first.
struct LanePipe {
    zkw_ctx* ctx;
    Transcript& tr;
    size_t n;
    struct Slot { int lane; bool ready, written; std::array<uint64_t, 8> xy; };
    std::vector<Slot> slots;     // by ticket
    std::vector<int> lane_owner; // lane -> ticket in flight, or -1
    LanePipe(zkw_ctx* c, Transcript& t, size_t n_) : ctx(c), tr(t), n(n_), lane_owner(zkw_ctx::kMsmLanes, -1) {}
    int retire(int ticket) {
        Slot& s = slots[ticket];
        if (s.ready) return ZKW_OK;
        uint64_t xyz[12];
        ZKW_TRY(msm_lane_wait(ctx, s.lane, xyz));
        memcpy(s.xy.data(), xyz, 64);
        bool ident = true;
        for (int i = 8; i < 12; i++) ident = ident && xyz[i] == 0;
        if (ident) s.xy.fill(0);
        s.ready = true;
        lane_owner[s.lane] = -1;
        return ZKW_OK;
    }
    int submit(int which, const uint64_t* poly, int* ticket = nullptr) {
        int lane = -1;
        for (int l = 1; l < zkw_ctx::kMsmLanes; l++) if (lane_owner[l] < 0) { lane = l; break; }
        if (lane < 0) {
            int oldest = -1;
            for (int l = 1; l < zkw_ctx::kMsmLanes; l++) if (oldest < 0 || lane_owner[l] < oldest) { oldest = lane_owner[l]; lane = l; }
            ZKW_TRY(retire(oldest));
        }
        ZKW_TRY(msm_lane_submit(ctx, lane, MsmJob{which, nullptr, poly, n}));
        slots.push_back(Slot{lane, false, false, {}});
        lane_owner[lane] = (int)slots.size() - 1;
        if (ticket) *ticket = (int)slots.size() - 1;
        return ZKW_OK;
    }
    int write(int ticket) {
        ZKW_TRY(retire(ticket));
        if (!slots[ticket].written) tr.write_point(slots[ticket].xy.data());
        slots[ticket].written = true;
        return ZKW_OK;
    }
    // every ticket not yet written, in submission order
    int flush() {
        for (int t = 0; t < (int)slots.size(); t++) if (!slots[t].written) ZKW_TRY(write(t));
        return ZKW_OK;
    }
};

// inclusive scan of x (n elements) into out; MUL / REVERSE as in prover_kernels.cuh; tmp holds block totals
template <bool MUL, bool REVERSE>
static int scan_run(zkw_ctx* ctx, const uint64_t* x, uint64_t* out, size_t n, uint64_t* tmp_blocks, uint64_t* grand_total_dev) {
    const size_t nblocks = (n + kScanBlock - 1) / kScanBlock;
    { ProfScope ps_(ctx, "scan_reduce_kernel"); scan_reduce_kernel<MUL, REVERSE><<<(unsigned)nblocks, kScanThreads, 0, ctx->stream>>>((const uint4*)x, (uint4*)tmp_blocks, n); }
    ZKW_LAUNCHED(ctx);
    { ProfScope ps_(ctx, "scan_blocks_kernel"); scan_blocks_kernel<MUL><<<1, 1024, 0, ctx->stream>>>((uint4*)tmp_blocks, nblocks, (uint4*)grand_total_dev); }
    ZKW_LAUNCHED(ctx);
    { ProfScope ps_(ctx, "scan_apply_kernel"); scan_apply_kernel<MUL, REVERSE><<<(unsigned)nblocks, kScanThreads, 0, ctx->stream>>>((const uint4*)x, (const uint4*)tmp_blocks, (uint4*)out, n); }
    ZKW_LAUNCHED(ctx);
    return ZKW_OK;
}

static int rand_fill(zkw_ctx* ctx, uint64_t* out, size_t count, const RandKey& seed, uint64_t stream, uint64_t first) {
    if (!count) return ZKW_OK;
    { ProfScope ps_(ctx, "rand_fill_kernel"); rand_fill_kernel<<<grid_for(count, 128), 128, 0, ctx->stream>>>((uint4*)out, count, seed, stream, first); }
    ZKW_LAUNCHED(ctx);
    return ZKW_OK;
}

// z = grand product over rows: z[0] = *z0 (or 1), z[r] = z0 prod_{i<r} num[i]/den[i] for r <= u, blinding above
static int grand_product(zkw_ctx* ctx, const uint64_t* num, const uint64_t* den, uint64_t* pn, uint64_t* sd, uint64_t* tmp_blocks,
                         uint64_t* total_dev, const uint64_t* z0_dev, uint64_t* z, size_t u, size_t n, const RandKey& seed, uint64_t stream) {
    ZKW_TRY((scan_run<true, false>(ctx, num, pn, n, tmp_blocks, nullptr)));
    ZKW_TRY((scan_run<true, true>(ctx, den, sd, n, tmp_blocks, total_dev)));
    Fr total;
    ZKW_CUDA(ctx, cudaMemcpyAsync(total.l, total_dev, 32, cudaMemcpyDeviceToHost, ctx->stream));
    ZKW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const Fr tinv = fp_inv_bingcd(total);
    { ProfScope ps_(ctx, "grand_product_finalize_kernel"); grand_product_finalize_kernel<<<grid_for(u + 1, 128), 128, 0, ctx->stream>>>((const uint4*)pn, (const uint4*)sd, (const uint4*)z0_dev, tinv, (uint4*)z, u, n); }
    ZKW_LAUNCHED(ctx);
    return rand_fill(ctx, z + 4 * (u + 1), n - (u + 1), seed, stream, 0);
}

struct Domain {
    unsigned k, ek;
    size_t n, en;
    DomainConsts dc;
    Fr omega, omega_inv;
    const uint64_t* tw = nullptr;  // omega^i, i < n/2 (device)
};

static int make_domain(zkw_ctx* ctx, const zkw_circuit_shape& sh, Domain* d) {
    d->k = sh.k; d->ek = sh.ext_k;
    d->n = (size_t)1 << sh.k; d->en = (size_t)1 << sh.ext_k;
    domain_consts(sh.k, sh.ext_k, &d->dc);
    d->omega = fr_of(d->dc.omega);
    d->omega_inv = fr_of(d->dc.omega_inv);
    return ntt_get_twiddles(ctx, d->dc.omega, sh.k, &d->tw);
}

static Fr rotate(const Domain& d, const Fr& x, int rot) {
    if (rot >= 0) return x * d.omega.pow((uint64_t)rot);
    return x * d.omega_inv.pow((uint64_t)(-rot));
}

}  // namespace zkw

extern "C" {

void zkw_pk_destroy(zkw_ctx* ctx, zkw_pk* pk) {
    if (!pk) return;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    for (void* p : pk->owned) cudaFree(p);
    delete pk;
}

// fixed_values: [F constants, table, A gate selectors, (q_lookup)] each n*4 u64 Montgomery (host);
// perm_mapping: [F + A + L] each n (col', row') u32 pairs (host) — halo2's keygen Assembly mapping.
// What both keygen and zkw_pk_read derive from the columns: l_0 / l_last / l_active on the extended coset and the
// sorted lookup table with multiplicities.  table_host: the table column's n values (Montgomery, host).
static int pk_derived(zkw_ctx* ctx, zkw_pk* pkp, const Domain& dom, const uint64_t* table_host) {
    struct Holder { zkw_pk* p; zkw_pk* get() const { return p; } zkw_pk* operator->() const { return p; } } pk{pkp};
    const zkw_circuit_shape& sh = pkp->shape;
    const size_t n = pkp->n, vb = n * 32, eb = pkp->en * 32;
    cudaStream_t st = ctx->stream;
    const uint64_t* const fixed_table = table_host;
    // l0, l_last, l_active on the extended coset
    {
        std::vector<uint64_t> h(n * 4, 0);
        Fr one = Fr::one();
        Scratch sc(ctx);
        uint64_t *d_v, *d_poly_unused;
        ZKW_TRY(sc.get(vb, (void**)&d_v));
        auto make = [&](uint64_t** coset) -> int {
            ZKW_CUDA(ctx, cudaMemcpyAsync(d_v, h.data(), vb, cudaMemcpyHostToDevice, st));
            ZKW_CUDA(ctx, cudaStreamSynchronize(st));
            ZKW_TRY(dmalloc(ctx, pk.get(), eb, (void**)coset));
            ZKW_TRY(ntt_run(ctx, d_v, sh.k, d_v, sh.k, dom.dc.omega_inv, false, dom.dc.n_scale3));
            return ntt_run(ctx, d_v, sh.k, *coset, sh.ext_k, dom.dc.ext_omega, true, nullptr);
        };
        (void)d_poly_unused;
        memcpy(&h[0], one.l, 32);
        ZKW_TRY(make(&pk->l0_coset));
        std::fill(h.begin(), h.end(), 0);
        memcpy(&h[4 * pk->u], one.l, 32);
        ZKW_TRY(make(&pk->l_last_coset));
        std::fill(h.begin(), h.end(), 0);
        for (size_t i = 0; i < pk->u; i++) memcpy(&h[4 * i], one.l, 32);
        ZKW_TRY(make(&pk->l_active_coset));
    }
    // lookup table: sorted distinct values over the usable rows, with multiplicities (host sort, keygen time)
    {
        const uint64_t* tab = fixed_table;
        std::vector<std::array<uint64_t, 4>> canon(pk->u);
        for (size_t i = 0; i < pk->u; i++) {
            Fr v; memcpy(v.l, tab + 4 * i, 32);
            v = v.from_mont();
            memcpy(canon[i].data(), v.l, 32);
        }
        auto less = [](const std::array<uint64_t, 4>& a, const std::array<uint64_t, 4>& b) {
            for (int i = 3; i >= 0; i--) if (a[i] != b[i]) return a[i] < b[i];
            return false;
        };
        std::sort(canon.begin(), canon.end(), less);
        std::vector<uint64_t> vals_c, vals_m;
        std::vector<uint32_t> mult;
        for (size_t i = 0; i < canon.size(); i++) {
            if (i && canon[i] == canon[i - 1]) { mult.back()++; continue; }
            vals_c.insert(vals_c.end(), canon[i].begin(), canon[i].end());
            Fr v; memcpy(v.l, canon[i].data(), 32);
            v = v.to_mont();
            uint64_t m4[4]; memcpy(m4, v.l, 32);
            vals_m.insert(vals_m.end(), m4, m4 + 4);
            mult.push_back(1);
        }
        pk->table_m = (uint32_t)mult.size();
        ZKW_TRY(dmalloc(ctx, pk.get(), vals_c.size() * 8, (void**)&pk->table_canon));
        ZKW_TRY(dmalloc(ctx, pk.get(), vals_m.size() * 8, (void**)&pk->table_mont));
        ZKW_TRY(dmalloc(ctx, pk.get(), mult.size() * 4, (void**)&pk->table_mult));
        ZKW_CUDA(ctx, cudaMemcpyAsync(pk->table_canon, vals_c.data(), vals_c.size() * 8, cudaMemcpyHostToDevice, st));
        ZKW_CUDA(ctx, cudaMemcpyAsync(pk->table_mont, vals_m.data(), vals_m.size() * 8, cudaMemcpyHostToDevice, st));
        ZKW_CUDA(ctx, cudaMemcpyAsync(pk->table_mult, mult.data(), mult.size() * 4, cudaMemcpyHostToDevice, st));
        ZKW_CUDA(ctx, cudaStreamSynchronize(st));
    }
    return ZKW_OK;
}

int zkw_keygen(zkw_ctx* ctx, const zkw_circuit_shape* shape, const uint64_t* const* fixed_values, const uint32_t* const* perm_mapping,
               zkw_pk** out) {
    if (!ctx || !shape || !fixed_values || !perm_mapping || !out) return ZKW_ERR_INVALID;
    ZKW_CUDA(ctx, cudaSetDevice(ctx->device));
    *out = nullptr;
    const zkw_circuit_shape& sh = *shape;
    if (sh.k < 4 || sh.k > 26 || sh.cs_degree < 4 || sh.cs_degree > 5 || sh.num_advice == 0 || sh.num_fixed == 0) return ZKW_ERR_INVALID;
    if (sh.num_lookup_advice == 0 && sh.num_advice != 1) return ZKW_ERR_UNSUPPORTED;
    if ((sh.num_lookup_advice == 0) != (sh.cs_degree == 5)) return ZKW_ERR_INVALID;
    {
        unsigned ek = sh.k;
        while ((1ull << ek) < (1ull << sh.k) * (sh.cs_degree - 1)) ek++;
        if (ek != sh.ext_k) return ZKW_ERR_INVALID;
    }
    if (!ctx->bases[ZKW_BASES_G].points || !ctx->bases[ZKW_BASES_G_LAGRANGE].points || ctx->bases[ZKW_BASES_G].n != ((size_t)1 << sh.k) ||
        ctx->bases[ZKW_BASES_G_LAGRANGE].n != ((size_t)1 << sh.k))
        return ZKW_ERR_STATE;  // the resident SRS (g and g_lagrange) must be the one for this domain size
    std::unique_ptr<zkw_pk, void (*)(zkw_pk*)> pk(new zkw_pk(), [](zkw_pk* p) { for (void* q : p->owned) cudaFree(q); delete p; });
    pk->shape = sh;
    pk->n = (size_t)1 << sh.k; pk->en = (size_t)1 << sh.ext_k;
    pk->u = pk->n - (sh.blinding_factors + 1);
    pk->A = sh.num_advice; pk->L = sh.num_lookup_advice; pk->F = sh.num_fixed;
    pk->nfixed = pk->F + 1 + pk->A + (pk->L == 0 ? 1 : 0);
    pk->nperm = pk->F + pk->A + pk->L;
    pk->chunk = sh.cs_degree - 2;
    pk->nsets = (pk->nperm + pk->chunk - 1) / pk->chunk;
    pk->nlk = pk->L ? pk->L : 1;
    if (pk->chunk > 8) return ZKW_ERR_UNSUPPORTED;
    const size_t n = pk->n, en = pk->en, vb = n * 32, eb = en * 32;
    Domain dom;
    ZKW_TRY(make_domain(ctx, sh, &dom));
    cudaStream_t st = ctx->stream;

    auto to_poly_and_coset = [&](uint64_t* values, uint64_t** poly, uint64_t** coset) -> int {
        ZKW_TRY(dmalloc(ctx, pk.get(), vb, (void**)poly));
        ZKW_TRY(dmalloc(ctx, pk.get(), eb, (void**)coset));
        ZKW_CUDA(ctx, cudaMemcpyAsync(*poly, values, vb, cudaMemcpyDeviceToDevice, st));
        ZKW_TRY(ntt_run(ctx, *poly, sh.k, *poly, sh.k, dom.dc.omega_inv, false, dom.dc.n_scale3));
        return ntt_run(ctx, *poly, sh.k, *coset, sh.ext_k, dom.dc.ext_omega, true, nullptr);
    };

    // fixed columns
    pk->fixed_values.resize(pk->nfixed); pk->fixed_polys.resize(pk->nfixed); pk->fixed_cosets.resize(pk->nfixed);
    pk->fixed_commitments.resize(pk->nfixed);
    for (unsigned c = 0; c < pk->nfixed; c++) {
        if (!fixed_values[c]) return ZKW_ERR_INVALID;
        ZKW_TRY(dmalloc(ctx, pk.get(), vb, (void**)&pk->fixed_values[c]));
        ZKW_CUDA(ctx, cudaMemcpyAsync(pk->fixed_values[c], fixed_values[c], vb, cudaMemcpyHostToDevice, st));
        ZKW_TRY(commit_dev(ctx, ZKW_BASES_G_LAGRANGE, pk->fixed_values[c], n, pk->fixed_commitments[c].data()));
        ZKW_TRY(to_poly_and_coset(pk->fixed_values[c], &pk->fixed_polys[c], &pk->fixed_cosets[c]));
    }
    // permutation: sigma values = delta^c' omega^r'
    {
        std::vector<Fr> dp(pk->nperm);
        Fr d = Fr::one(), delta = fr_of(kDeltaM);
        for (unsigned c = 0; c < pk->nperm; c++) { dp[c] = d; d = d * delta; }
        Scratch sc(ctx);
        uint64_t* d_dp; uint2* d_map;
        ZKW_TRY(sc.get(pk->nperm * 32, (void**)&d_dp));
        ZKW_TRY(sc.get(n * 8, (void**)&d_map));
        ZKW_CUDA(ctx, cudaMemcpyAsync(d_dp, dp.data(), pk->nperm * 32, cudaMemcpyHostToDevice, st));
        pk->sigma_values.resize(pk->nperm); pk->sigma_polys.resize(pk->nperm); pk->sigma_cosets.resize(pk->nperm);
        pk->perm_commitments.resize(pk->nperm);
        for (unsigned c = 0; c < pk->nperm; c++) {
            if (!perm_mapping[c]) return ZKW_ERR_INVALID;
            for (size_t i = 0; i < n; i++)
                if (perm_mapping[c][2 * i] >= pk->nperm || perm_mapping[c][2 * i + 1] >= n) return ZKW_ERR_INVALID;
            ZKW_CUDA(ctx, cudaMemcpyAsync(d_map, perm_mapping[c], n * 8, cudaMemcpyHostToDevice, st));
            ZKW_TRY(dmalloc(ctx, pk.get(), vb, (void**)&pk->sigma_values[c]));
            { ProfScope ps_(ctx, "sigma_values_kernel"); sigma_values_kernel<<<grid_for(n, 128), 128, 0, st>>>(d_map, (const uint4*)d_dp, (const uint4*)dom.tw, (uint4*)pk->sigma_values[c], n); }
            ZKW_LAUNCHED(ctx);
            ZKW_TRY(commit_dev(ctx, ZKW_BASES_G_LAGRANGE, pk->sigma_values[c], n, pk->perm_commitments[c].data()));
            ZKW_TRY(to_poly_and_coset(pk->sigma_values[c], &pk->sigma_polys[c], &pk->sigma_cosets[c]));
        }
    }
    ZKW_TRY(pk_derived(ctx, pk.get(), dom, fixed_values[pk->table_col()]));
    // VK digest (self-defined, see oracle/halo2_ref.py::vk_digest): Blake2b-512("Halo2-Verify-Key") mod r
    {
        Blake2b h(64, "Halo2-Verify-Key");
        h.update("zkw-b200-vk-v1", 14);
        const uint32_t hdr[6] = {sh.k, sh.num_advice, sh.num_lookup_advice, sh.num_fixed, sh.blinding_factors, sh.cs_degree};
        h.update(hdr, 24);
        auto add_pt = [&](const std::array<uint64_t, 8>& p) {
            Fq x, y; memcpy(x.l, p.data(), 32); memcpy(y.l, p.data() + 4, 32);
            x = x.from_mont(); y = y.from_mont();
            h.update(x.l, 32); h.update(y.l, 32);
        };
        for (auto& p : pk->fixed_commitments) add_pt(p);
        for (auto& p : pk->perm_commitments) add_pt(p);
        uint8_t dg[64];
        h.peek_digest(dg);
        Fr lo, hi; memcpy(lo.l, dg, 32); memcpy(hi.l, dg + 32, 32);
        Fr d = lo * Fr::r2() + (hi * Fr::r2()) * Fr::r2();
        memcpy(pk->digest, d.l, 32);
    }
    ZKW_CUDA(ctx, cudaStreamSynchronize(st));
    *out = pk.release();
    return ZKW_OK;
}

int zkw_pk_info(const zkw_pk* pk, uint32_t* num_fixed_cols, uint32_t* num_perm_cols) {
    if (!pk) return ZKW_ERR_INVALID;
    if (num_fixed_cols) *num_fixed_cols = pk->nfixed;
    if (num_perm_cols) *num_perm_cols = pk->nperm;
    return ZKW_OK;
}

// Verifying key: commitments (affine Montgomery, 8 u64 each) and the transcript digest (Montgomery Fr).
int zkw_pk_vk(const zkw_pk* pk, uint64_t* fixed_commitments_xy, uint64_t* perm_commitments_xy, uint64_t digest[4]) {
    if (!pk) return ZKW_ERR_INVALID;
    if (fixed_commitments_xy) for (unsigned c = 0; c < pk->nfixed; c++) memcpy(fixed_commitments_xy + 8 * c, pk->fixed_commitments[c].data(), 64);
    if (perm_commitments_xy) for (unsigned c = 0; c < pk->nperm; c++) memcpy(perm_commitments_xy + 8 * c, pk->perm_commitments[c].data(), 64);
    if (digest) memcpy(digest, pk->digest, 32);
    return ZKW_OK;
}

// ---- key files: the replacement of ProvingKey / VerifyingKey ::to_bytes / ::read with SerdeFormat::RawBytes -------
// (ecdsa_p256.rs:261-270 writes them in download_keys, :339-343 / :388-393 read the proving key on every request).
// "Raw bytes" = the in-memory Montgomery limbs, as upstream's RawBytes.  Layout (little-endian):
//   magic "ZKWPK1\0\0" | zkw_circuit_shape (8 u32) | nfixed u32 | nperm u32 | digest 32 B
//   | fixed commitments nfixed x 64 B | permutation commitments nperm x 64 B            <- the verifying key ends here
//   | per fixed column: values n x 32 B, polynomial n x 32 B | per permutation column: sigma values, sigma polynomial
// Extended cosets, l_0 / l_last / l_active and the sorted lookup table are NOT stored (upstream stores the cosets):
// they are one NTT each on load, which is faster than reading 4n more field elements per column from disk.
static const char kPkMagic[8] = {'Z', 'K', 'W', 'P', 'K', '1', 0, 0};
static const char kVkMagic[8] = {'Z', 'K', 'W', 'V', 'K', '1', 0, 0};

static int write_vk_part(FILE* f, const zkw_pk* pk, const char magic[8]) {
    const uint32_t counts[2] = {pk->nfixed, pk->nperm};
    if (fwrite(magic, 1, 8, f) != 8 || fwrite(&pk->shape, sizeof(zkw_circuit_shape), 1, f) != 1 || fwrite(counts, 4, 2, f) != 2 ||
        fwrite(pk->digest, 8, 4, f) != 4)
        return ZKW_ERR_INVALID;
    for (auto& c : pk->fixed_commitments) if (fwrite(c.data(), 8, 8, f) != 8) return ZKW_ERR_INVALID;
    for (auto& c : pk->perm_commitments) if (fwrite(c.data(), 8, 8, f) != 8) return ZKW_ERR_INVALID;
    return ZKW_OK;
}

int zkw_vk_write(const zkw_pk* pk, const char* path) {
    if (!pk || !path) return ZKW_ERR_INVALID;
    FILE* f = fopen(path, "wb");
    if (!f) return ZKW_ERR_INVALID;
    int rc = write_vk_part(f, pk, kVkMagic);
    if (fclose(f) != 0) rc = ZKW_ERR_INVALID;
    return rc;
}

int zkw_vk_read(const char* path, zkw_circuit_shape* shape, uint32_t* num_fixed_cols, uint32_t* num_perm_cols, uint64_t* fixed_commitments_xy,
                size_t fixed_cap, uint64_t* perm_commitments_xy, size_t perm_cap, uint64_t digest[4]) {
    if (!path) return ZKW_ERR_INVALID;
    FILE* f = fopen(path, "rb");
    if (!f) return ZKW_ERR_INVALID;
    char magic[8];
    zkw_circuit_shape sh;
    uint32_t counts[2];
    uint64_t dg[4];
    int rc = ZKW_OK;
    if (fread(magic, 1, 8, f) != 8 || (memcmp(magic, kVkMagic, 8) && memcmp(magic, kPkMagic, 8)) || fread(&sh, sizeof sh, 1, f) != 1 ||
        fread(counts, 4, 2, f) != 2 || fread(dg, 8, 4, f) != 4)
        rc = ZKW_ERR_INVALID;
    if (rc == ZKW_OK) {
        if (shape) *shape = sh;
        if (num_fixed_cols) *num_fixed_cols = counts[0];
        if (num_perm_cols) *num_perm_cols = counts[1];
        if (digest) memcpy(digest, dg, 32);
        std::vector<uint64_t> pts(8 * ((size_t)counts[0] + counts[1]));
        if (counts[0] > 4096 || counts[1] > 4096 || fread(pts.data(), 8, pts.size(), f) != pts.size()) rc = ZKW_ERR_INVALID;
        else {
            if (fixed_commitments_xy) { if (fixed_cap < counts[0]) rc = ZKW_ERR_INVALID; else memcpy(fixed_commitments_xy, pts.data(), 64 * (size_t)counts[0]); }
            if (perm_commitments_xy) { if (perm_cap < counts[1]) rc = ZKW_ERR_INVALID; else memcpy(perm_commitments_xy, pts.data() + 8 * (size_t)counts[0], 64 * (size_t)counts[1]); }
        }
    }
    fclose(f);
    return rc;
}

int zkw_pk_write(zkw_ctx* ctx, const zkw_pk* pk, const char* path) {
    if (!ctx || !pk || !path) return ZKW_ERR_INVALID;
    ZKW_CUDA(ctx, cudaSetDevice(ctx->device));
    FILE* f = fopen(path, "wb");
    if (!f) return ZKW_ERR_INVALID;
    int rc = write_vk_part(f, pk, kPkMagic);
    const size_t vb = pk->n * 32;
    std::vector<uint8_t> host(vb);
    auto dump = [&](const uint64_t* dev) -> int {
        if (cudaMemcpy(host.data(), dev, vb, cudaMemcpyDeviceToHost) != cudaSuccess) return ZKW_ERR_CUDA;
        return fwrite(host.data(), 1, vb, f) == vb ? ZKW_OK : ZKW_ERR_INVALID;
    };
    for (unsigned c = 0; c < pk->nfixed && rc == ZKW_OK; c++) { rc = dump(pk->fixed_values[c]); if (rc == ZKW_OK) rc = dump(pk->fixed_polys[c]); }
    for (unsigned c = 0; c < pk->nperm && rc == ZKW_OK; c++) { rc = dump(pk->sigma_values[c]); if (rc == ZKW_OK) rc = dump(pk->sigma_polys[c]); }
    if (fclose(f) != 0 && rc == ZKW_OK) rc = ZKW_ERR_INVALID;
    return rc;
}

int zkw_pk_read(zkw_ctx* ctx, const char* path, zkw_pk** out) {
    if (!ctx || !path || !out) return ZKW_ERR_INVALID;
    *out = nullptr;
    ZKW_CUDA(ctx, cudaSetDevice(ctx->device));
    FILE* f = fopen(path, "rb");
    if (!f) return ZKW_ERR_INVALID;
    std::unique_ptr<FILE, int (*)(FILE*)> closer(f, fclose);
    char magic[8];
    zkw_circuit_shape sh;
    uint32_t counts[2];
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, kPkMagic, 8) || fread(&sh, sizeof sh, 1, f) != 1 || fread(counts, 4, 2, f) != 2) return ZKW_ERR_INVALID;
    if (sh.k < 4 || sh.k > 26 || sh.cs_degree < 4 || sh.cs_degree > 5 || sh.num_advice == 0 || sh.num_fixed == 0 || sh.ext_k < sh.k || sh.ext_k > sh.k + 2)
        return ZKW_ERR_INVALID;
    if (!ctx->bases[ZKW_BASES_G].points || ctx->bases[ZKW_BASES_G].n != ((size_t)1 << sh.k)) return ZKW_ERR_STATE;
    std::unique_ptr<zkw_pk, void (*)(zkw_pk*)> pk(new zkw_pk(), [](zkw_pk* p) { for (void* q : p->owned) cudaFree(q); delete p; });
    pk->shape = sh;
    pk->n = (size_t)1 << sh.k; pk->en = (size_t)1 << sh.ext_k;
    pk->u = pk->n - (sh.blinding_factors + 1);
    pk->A = sh.num_advice; pk->L = sh.num_lookup_advice; pk->F = sh.num_fixed;
    pk->nfixed = pk->F + 1 + pk->A + (pk->L == 0 ? 1 : 0);
    pk->nperm = pk->F + pk->A + pk->L;
    pk->chunk = sh.cs_degree - 2;
    pk->nsets = (pk->nperm + pk->chunk - 1) / pk->chunk;
    pk->nlk = pk->L ? pk->L : 1;
    if (counts[0] != pk->nfixed || counts[1] != pk->nperm || pk->chunk > 8) return ZKW_ERR_INVALID;
    if (fread(pk->digest, 8, 4, f) != 4) return ZKW_ERR_INVALID;
    pk->fixed_commitments.resize(pk->nfixed); pk->perm_commitments.resize(pk->nperm);
    for (auto& c : pk->fixed_commitments) if (fread(c.data(), 8, 8, f) != 8) return ZKW_ERR_INVALID;
    for (auto& c : pk->perm_commitments) if (fread(c.data(), 8, 8, f) != 8) return ZKW_ERR_INVALID;
    const size_t n = pk->n, vb = n * 32, eb = pk->en * 32;
    Domain dom;
    ZKW_TRY(make_domain(ctx, sh, &dom));
    cudaStream_t st = ctx->stream;
    std::vector<uint64_t> host(4 * n), table_host;
    auto load = [&](uint64_t** dev, bool keep_table) -> int {
        if (fread(host.data(), 1, vb, f) != vb) return ZKW_ERR_INVALID;
        if (keep_table) table_host = host;
        ZKW_TRY(dmalloc(ctx, pk.get(), vb, (void**)dev));
        ZKW_CUDA(ctx, cudaMemcpyAsync(*dev, host.data(), vb, cudaMemcpyHostToDevice, st));
        ZKW_CUDA(ctx, cudaStreamSynchronize(st));     // `host` is reused by the next column
        return ZKW_OK;
    };
    auto coset_of = [&](const uint64_t* poly, uint64_t** coset) -> int {
        ZKW_TRY(dmalloc(ctx, pk.get(), eb, (void**)coset));
        return ntt_run(ctx, poly, sh.k, *coset, sh.ext_k, dom.dc.ext_omega, true, nullptr);
    };
    pk->fixed_values.resize(pk->nfixed); pk->fixed_polys.resize(pk->nfixed); pk->fixed_cosets.resize(pk->nfixed);
    for (unsigned c = 0; c < pk->nfixed; c++) {
        ZKW_TRY(load(&pk->fixed_values[c], c == pk->table_col()));
        ZKW_TRY(load(&pk->fixed_polys[c], false));
        ZKW_TRY(coset_of(pk->fixed_polys[c], &pk->fixed_cosets[c]));
    }
    pk->sigma_values.resize(pk->nperm); pk->sigma_polys.resize(pk->nperm); pk->sigma_cosets.resize(pk->nperm);
    for (unsigned c = 0; c < pk->nperm; c++) {
        ZKW_TRY(load(&pk->sigma_values[c], false));
        ZKW_TRY(load(&pk->sigma_polys[c], false));
        ZKW_TRY(coset_of(pk->sigma_polys[c], &pk->sigma_cosets[c]));
    }
    ZKW_TRY(pk_derived(ctx, pk.get(), dom, table_host.data()));
    ZKW_CUDA(ctx, cudaStreamSynchronize(st));
    *out = pk.release();
    return ZKW_OK;
}

// advice: [A + L] host arrays of advice_rows[c] <= usable rows field elements (Montgomery); the remaining
// usable rows are zero (unassigned cells), the last blinding_factors+1 rows are blinding.
// transcript: 0 = Blake2b/Challenge255 (compressed points), 1 = EVM/keccak (uncompressed).
static int create_proof_impl(zkw_ctx* ctx, const zkw_pk* pk, const uint64_t* const* advice, const size_t* advice_rows, const RandKey& seed,
                             int transcript, unsigned flags, uint8_t* out, size_t out_cap, size_t* out_len, zkw_advice_ready_fn ready = nullptr,
                             void* ready_user = nullptr);

// 64-bit seeds (deterministic streams for tests and A/B runs) are widened with zeros; production callers pass 32
// bytes from the OS through zkw_create_proof_seeded
int zkw_create_proof_ex(zkw_ctx* ctx, const zkw_pk* pk, const uint64_t* const* advice, const size_t* advice_rows, uint64_t seed,
                        int transcript, unsigned flags, uint8_t* out, size_t out_cap, size_t* out_len) {
    RandKey key = {{(uint32_t)seed, (uint32_t)(seed >> 32), 0, 0, 0, 0, 0, 0}};
    return create_proof_impl(ctx, pk, advice, advice_rows, key, transcript, flags, out, out_cap, out_len);
}

int zkw_create_proof(zkw_ctx* ctx, const zkw_pk* pk, const uint64_t* const* advice, const size_t* advice_rows, uint64_t seed,
                     int transcript, uint8_t* out, size_t out_cap, size_t* out_len) {
    return zkw_create_proof_ex(ctx, pk, advice, advice_rows, seed, transcript, 0u, out, out_cap, out_len);
}

int zkw_create_proof_seeded(zkw_ctx* ctx, const zkw_pk* pk, const uint64_t* const* advice, const size_t* advice_rows, const uint8_t seed[32],
                            int transcript, unsigned flags, uint8_t* out, size_t out_cap, size_t* out_len) {
    if (!seed) return ZKW_ERR_INVALID;
    RandKey key;
    memcpy(key.w, seed, 32);
    return create_proof_impl(ctx, pk, advice, advice_rows, key, transcript, flags, out, out_cap, out_len);
}

int zkw_create_proof_overlapped(zkw_ctx* ctx, const zkw_pk* pk, const uint64_t* const* advice, const size_t* advice_rows, const uint8_t seed[32],
                                int transcript, unsigned flags, zkw_advice_ready_fn ready, void* user, uint8_t* out, size_t out_cap, size_t* out_len) {
    if (!seed) return ZKW_ERR_INVALID;
    RandKey key;
    memcpy(key.w, seed, 32);
    return create_proof_impl(ctx, pk, advice, advice_rows, key, transcript, flags, out, out_cap, out_len, ready, user);
}

static int create_proof_impl(zkw_ctx* ctx, const zkw_pk* pk, const uint64_t* const* advice, const size_t* advice_rows, const RandKey& seed,
                             int transcript, unsigned flags, uint8_t* out, size_t out_cap, size_t* out_len, zkw_advice_ready_fn ready,
                             void* ready_user) {
    if (!ctx || !pk || !advice || !advice_rows || !out_len || (transcript != 0 && transcript != 1)) return ZKW_ERR_INVALID;
    const bool adv_on_device = flags & ZKW_ADVICE_ON_DEVICE, adv_canonical = flags & ZKW_ADVICE_CANONICAL, shplonk = flags & ZKW_MULTIOPEN_SHPLONK,
               adv_u64 = flags & ZKW_ADVICE_U64;
    ZKW_CUDA(ctx, cudaSetDevice(ctx->device));
    const zkw_circuit_shape& sh = pk->shape;
    const size_t n = pk->n, en = pk->en, u = pk->u, vb = n * 32, eb = en * 32;
    const unsigned A = pk->A, L = pk->L, F = pk->F, NA = A + L, nsets = pk->nsets, nlk = pk->nlk;
    if (ctx->bases[ZKW_BASES_G].n != n || ctx->bases[ZKW_BASES_G_LAGRANGE].n != n || !ctx->bases[ZKW_BASES_G_LAGRANGE].points) return ZKW_ERR_STATE;
    for (unsigned c = 0; c < NA; c++) if (advice_rows[c] > u || (advice_rows[c] && !advice[c])) return ZKW_ERR_INVALID;
    cudaStream_t st = ctx->stream;
    Domain dom;
    ZKW_TRY(make_domain(ctx, sh, &dom));
    Scratch sc(ctx);
    Transcript tr(transcript);
    tr.common_scalar(fr_of(pk->digest));

    // Auxiliary streams: as soon as a committed column is final, its coefficient form (iNTT) and extended
    // coset (zeta-coset NTT) are computed there, overlapping the commitments running on the main stream
    // and the MSM lanes.  ZKW_AUX_STREAMS=2|3 spreads the columns round robin over more streams so that independent
    // transforms overlap each other as well; measured (k = 17 / 18 / 19 resident proofs, tools/proof_ab.py): 12.07 / 17.63 /
    // 26.54 ms with one stream, 12.16 / 18.02 / 26.59 with two, 12.18 / 17.95 / 26.68 with three — the device is already
    // multiplier-bound with the MSM lanes next to one transform stream, more concurrency only adds contention.  Default 1.
    // The main stream joins them all before the quotient kernel.
    if (!ctx->aux_count) {
        const char* e = getenv("ZKW_AUX_STREAMS");
        int want = e ? atoi(e) : 1;
        want = want < 1 ? 1 : want > zkw_ctx::kAuxStreams ? zkw_ctx::kAuxStreams : want;
        ZKW_CUDA(ctx, cudaEventCreateWithFlags(&ctx->aux_fork, cudaEventDisableTiming));
        for (int i = 0; i < want; i++) {
            ZKW_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->aux_stream[i], cudaStreamNonBlocking, stream_priority(1)));
            ZKW_CUDA(ctx, cudaEventCreateWithFlags(&ctx->aux_join[i], cudaEventDisableTiming));
        }
        ctx->aux_count = want;
    }
    auto spawn_transform = [&](const uint64_t* lagrange, uint64_t** coeff_out, const uint64_t** ext_out) -> int {
        uint64_t *cf, *ex;
        ZKW_TRY(sc.get(vb, (void**)&cf));
        ZKW_TRY(sc.get(eb, (void**)&ex));
        cudaStream_t sx = ctx->aux_stream[ctx->aux_next];
        ctx->aux_next = (ctx->aux_next + 1) % ctx->aux_count;
        ZKW_CUDA(ctx, cudaEventRecord(ctx->aux_fork, st));
        ZKW_CUDA(ctx, cudaStreamWaitEvent(sx, ctx->aux_fork, 0));
        ZKW_TRY(ntt_run(ctx, lagrange, sh.k, cf, sh.k, dom.dc.omega_inv, false, dom.dc.n_scale3, sx));
        ZKW_TRY(ntt_run(ctx, cf, sh.k, ex, sh.ext_k, dom.dc.ext_omega, true, nullptr, sx));
        *coeff_out = cf;
        *ext_out = ex;
        return ZKW_OK;
    };
    std::vector<uint64_t*> adv_cf(NA), perm_z_cf(nsets), lk_z_cf(nlk), lk_a_cf(nlk), lk_s_cf(nlk);
    std::vector<const uint64_t*> e_adv(NA), e_pz(nsets), e_lz(nlk), e_la(nlk), e_ls(nlk);

    // ---- 0. witness-independent work first: the vanishing argument's random polynomial and its commitment (absorbed by
    // the transcript after the grand products).  With zkw_create_proof_overlapped the caller is still synthesising the
    // witness on the host while this MSM runs; `ready` blocks until the advice columns may be read.
    const int lookup_probe = getenv("ZKW_LOOKUP_NO_PROBE") ? 0 : 1;   // test knob: force the binary search of lookup_rank_kernel
    LanePipe pipe(ctx, tr, n);
    uint64_t* random_poly;
    int t_random;
    ZKW_TRY(sc.get(vb, (void**)&random_poly));
    ZKW_TRY(rand_fill(ctx, random_poly, n, seed, 4000, 0));
    // Its MSM is a full one-wave accumulation: whatever is queued on the other lanes while it is resident waits for it
    // (a lane's one-CTA scan cannot get a slot).  When the witness is still being synthesised that is free time and the MSM
    // goes first; when the advice is already here it is submitted after the (light) advice and permuted-lookup commitments,
    // which the transcript needs first.
    const bool random_first = ready != nullptr;
    // ZKW_E2E_TRACE=1 (development aid): host clock at the hand-over points of the overlapped path, on stderr
    static const bool e2e_trace = getenv("ZKW_E2E_TRACE") != nullptr;
    timespec tr0{};
    auto trace = [&](const char* what) {
        if (!e2e_trace) return;
        timespec t; clock_gettime(CLOCK_MONOTONIC, &t);
        fprintf(stderr, "  [e2e] %-28s %.3f ms\n", what, 1e3 * (double)(t.tv_sec - tr0.tv_sec) + 1e-6 * (double)(t.tv_nsec - tr0.tv_nsec));
    };
    if (e2e_trace) clock_gettime(CLOCK_MONOTONIC, &tr0);
    if (random_first) ZKW_TRY(pipe.submit(ZKW_BASES_G, random_poly, &t_random));
    trace("random MSM submitted");
    if (ready) ZKW_TRY(ready(ready_user));
    trace("advice ready");

    // ---- 1. advice ----
    std::vector<uint64_t*> adv(NA);
    for (unsigned c = 0; c < NA; c++) {
        ZKW_TRY(sc.get(vb, (void**)&adv[c]));
        if (advice_rows[c] && adv_u64) {
            // compact form: one u64 per row, staged next to the column and widened on the device
            uint64_t* stage;
            ZKW_TRY(sc.get(advice_rows[c] * 8, (void**)&stage));
            ZKW_CUDA(ctx, cudaMemcpyAsync(stage, advice[c], advice_rows[c] * 8, adv_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
            { ProfScope ps_(ctx, "u64_to_mont_kernel"); u64_to_mont_kernel<<<grid_for(advice_rows[c], 128), 128, 0, st>>>(stage, (uint4*)adv[c], advice_rows[c]); }
            ZKW_LAUNCHED(ctx);
        } else if (advice_rows[c]) {
            ZKW_CUDA(ctx, cudaMemcpyAsync(adv[c], advice[c], advice_rows[c] * 32, adv_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
            if (adv_canonical) {
                { ProfScope ps_(ctx, "to_mont_kernel"); to_mont_kernel<<<grid_for(advice_rows[c], 128), 128, 0, st>>>((uint4*)adv[c], advice_rows[c]); }
                ZKW_LAUNCHED(ctx);
            }
        }
        if (u > advice_rows[c]) {
            { ProfScope ps_(ctx, "zero_fill_kernel"); zero_fill_kernel<<<grid_for(u - advice_rows[c], 128), 128, 0, st>>>((uint4*)(adv[c] + 4 * advice_rows[c]), u - advice_rows[c]); }
            ZKW_LAUNCHED(ctx);
        }
        ZKW_TRY(rand_fill(ctx, adv[c] + 4 * u, n - u, seed, 1 + c, 0));
        ZKW_TRY(spawn_transform(adv[c], &adv_cf[c], &e_adv[c]));
    }
    // Commitments of this and the next two rounds that do not depend on a challenge are all started now, each on
    // its own MSM lane: the advice columns, the permuted lookup columns (single-expression lookups: the theta
    // compression is the identity, so A' and S' are functions of the witness alone) and the random polynomial of
    // the vanishing argument.  The transcript still absorbs them in upstream's order (pipe.write below).
    std::vector<int> t_adv(NA), t_lk;
    trace("advice copies enqueued");
    if (e2e_trace) { cudaStreamSynchronize(st); trace("advice on the device"); }
    for (unsigned c = 0; c < NA; c++) ZKW_TRY(pipe.submit(ZKW_BASES_G_LAGRANGE, adv[c], &t_adv[c]));

    // ---- 2. lookups: permuted input / table ----
    uint64_t *num, *den, *pn, *sd, *blocks, *total_dev;
    ZKW_TRY(sc.get(vb, (void**)&num)); ZKW_TRY(sc.get(vb, (void**)&den));
    ZKW_TRY(sc.get(vb, (void**)&pn)); ZKW_TRY(sc.get(vb, (void**)&sd));
    ZKW_TRY(sc.get(((n + kScanBlock - 1) / kScanBlock + 1) * 32, (void**)&blocks));
    ZKW_TRY(sc.get(32, (void**)&total_dev));
    std::vector<uint64_t*> lk_inp(nlk), lk_a(nlk), lk_s(nlk), lk_z(nlk);
    {
        const uint32_t m = pk->table_m;
        uint32_t *rank, *counts, *run_start, *rep_start, *desc_start, *err;
        ZKW_TRY(sc.get(u * 4, (void**)&rank)); ZKW_TRY(sc.get((m + 1) * 4, (void**)&counts));
        ZKW_TRY(sc.get((m + 1) * 4, (void**)&run_start)); ZKW_TRY(sc.get((m + 1) * 4, (void**)&rep_start));
        ZKW_TRY(sc.get((m + 1) * 4, (void**)&desc_start)); ZKW_TRY(sc.get(4, (void**)&err));
        uint32_t* tile_tot;
        ZKW_TRY(sc.get(3 * std::max<size_t>(1, ((size_t)m + kLookupTile - 1) / kLookupTile) * 4, (void**)&tile_tot));
        ZKW_CUDA(ctx, cudaMemsetAsync(err, 0, 4, st));
        for (unsigned l = 0; l < nlk; l++) {
            if (L) lk_inp[l] = adv[A + l];
            else {
                ZKW_TRY(sc.get(vb, (void**)&lk_inp[l]));
                { ProfScope ps_(ctx, "mul_vec_kernel"); mul_vec_kernel<<<grid_for(n, 128), 128, 0, st>>>((const uint4*)pk->fixed_values[pk->q_lookup_col()], (const uint4*)adv[0], (uint4*)lk_inp[l], n); }
                ZKW_LAUNCHED(ctx);
            }
            ZKW_TRY(sc.get(vb, (void**)&lk_a[l])); ZKW_TRY(sc.get(vb, (void**)&lk_s[l])); ZKW_TRY(sc.get(vb, (void**)&lk_z[l]));
            ZKW_CUDA(ctx, cudaMemsetAsync(counts, 0, (m + 1) * 4, st));
            { ProfScope ps_(ctx, "lookup_rank_kernel"); lookup_rank_kernel<<<grid_for(u, 128), 128, 0, st>>>((const uint4*)lk_inp[l], (const uint4*)pk->table_canon, m, rank, counts, err, u, lookup_probe); }
            ZKW_LAUNCHED(ctx);
            {
                const unsigned tiles = std::max(1u, (m + kLookupTile - 1) / kLookupTile);
                { ProfScope ps_(ctx, "lookup_scan_totals_kernel"); lookup_scan_totals_kernel<<<tiles, 1024, 0, st>>>(counts, pk->table_mult, m, tile_tot, err); }
                ZKW_LAUNCHED(ctx);
                { ProfScope ps_(ctx, "lookup_scan_kernel"); lookup_scan_kernel<<<tiles, 1024, 0, st>>>(counts, pk->table_mult, m, tile_tot, run_start, rep_start, desc_start, err); }
            }
            ZKW_LAUNCHED(ctx);
            uint32_t herr = 0;
            ZKW_CUDA(ctx, cudaMemcpyAsync(&herr, err, 4, cudaMemcpyDeviceToHost, st));
            ZKW_CUDA(ctx, cudaStreamSynchronize(st));
            if (herr) return ZKW_ERR_INVALID;  // upstream: Error::ConstraintSystemFailure (lookup input not in table)
            { ProfScope ps_(ctx, "lookup_expand_kernel"); lookup_expand_kernel<<<grid_for(u, 128), 128, 0, st>>>((const uint4*)pk->table_mont, m, run_start, rep_start, desc_start, (uint4*)lk_a[l], (uint4*)lk_s[l], u); }
            ZKW_LAUNCHED(ctx);
            ZKW_TRY(rand_fill(ctx, lk_a[l] + 4 * u, n - u, seed, 1000 + 2 * l, 0));
            ZKW_TRY(rand_fill(ctx, lk_s[l] + 4 * u, n - u, seed, 1001 + 2 * l, 0));
            ZKW_TRY(spawn_transform(lk_a[l], &lk_a_cf[l], &e_la[l]));
            ZKW_TRY(spawn_transform(lk_s[l], &lk_s_cf[l], &e_ls[l]));
        }
        for (unsigned l = 0; l < nlk; l++) {
            int ta, ts;
            ZKW_TRY(pipe.submit(ZKW_BASES_G_LAGRANGE, lk_a[l], &ta));
            ZKW_TRY(pipe.submit(ZKW_BASES_G_LAGRANGE, lk_s[l], &ts));
            t_lk.push_back(ta); t_lk.push_back(ts);
        }
    }
    if (!random_first) ZKW_TRY(pipe.submit(ZKW_BASES_G, random_poly, &t_random));
    for (int t : t_adv) ZKW_TRY(pipe.write(t));
    const Fr theta = tr.squeeze();
    (void)theta;
    for (int t : t_lk) ZKW_TRY(pipe.write(t));
    const Fr beta = tr.squeeze();
    const Fr gamma = tr.squeeze();

    // ---- 3. permutation grand products ----
    std::vector<uint64_t*> perm_z(nsets);
    std::vector<int> t_z;
    {
        const Fr delta = fr_of(kDeltaM);
        Fr dpow = Fr::one();
        for (unsigned s = 0; s < nsets; s++) {
            ZKW_TRY(sc.get(vb, (void**)&perm_z[s]));
            PermChunkArgs a;
            memset(&a, 0, sizeof(a));
            const unsigned c0 = s * pk->chunk, c1 = std::min(c0 + pk->chunk, pk->nperm);
            a.ncols = (int)(c1 - c0);
            for (unsigned c = c0; c < c1; c++) {
                a.values[c - c0] = (const uint4*)(c < F ? pk->fixed_values[c] : adv[c - F]);
                a.sigmas[c - c0] = (const uint4*)pk->sigma_values[c];
                a.delta_beta[c - c0] = dpow * beta;
                dpow = dpow * delta;
            }
            a.beta = beta; a.gamma = gamma; a.tw = (const uint4*)dom.tw; a.n = n;
            { ProfScope ps_(ctx, "perm_numden_kernel"); perm_numden_kernel<<<grid_for(n, 128), 128, 0, st>>>(a, (uint4*)num, (uint4*)den); }
            ZKW_LAUNCHED(ctx);
            const uint64_t* z0 = s ? perm_z[s - 1] + 4 * u : nullptr;
            ZKW_TRY(grand_product(ctx, num, den, pn, sd, blocks, total_dev, z0, perm_z[s], u, n, seed, 2000 + s));
            { int t; ZKW_TRY(pipe.submit(ZKW_BASES_G_LAGRANGE, perm_z[s], &t)); t_z.push_back(t); }
            ZKW_TRY(spawn_transform(perm_z[s], &perm_z_cf[s], &e_pz[s]));
        }
    }
    // ---- 4. lookup grand products ----
    for (unsigned l = 0; l < nlk; l++) {
        { ProfScope ps_(ctx, "lookup_numden_kernel"); lookup_numden_kernel<<<grid_for(n, 128), 128, 0, st>>>((const uint4*)lk_inp[l], (const uint4*)pk->fixed_values[pk->table_col()], (const uint4*)lk_a[l], (const uint4*)lk_s[l], beta, gamma, (uint4*)num, (uint4*)den, n); }
        ZKW_LAUNCHED(ctx);
        ZKW_TRY(grand_product(ctx, num, den, pn, sd, blocks, total_dev, nullptr, lk_z[l], u, n, seed, 3000 + l));
        { int t; ZKW_TRY(pipe.submit(ZKW_BASES_G_LAGRANGE, lk_z[l], &t)); t_z.push_back(t); }
        ZKW_TRY(spawn_transform(lk_z[l], &lk_z_cf[l], &e_lz[l]));
    }
    // ---- 5. transcript: grand products, then the random polynomial's commitment ----
    for (int t : t_z) ZKW_TRY(pipe.write(t));
    ZKW_TRY(pipe.write(t_random));
    const Fr y = tr.squeeze();

    // ---- 6. quotient ----
    // the coefficient forms and extended cosets were produced on the auxiliary stream: join it
    for (int i = 0; i < ctx->aux_count; i++) {
        ZKW_CUDA(ctx, cudaEventRecord(ctx->aux_join[i], ctx->aux_stream[i]));
        ZKW_CUDA(ctx, cudaStreamWaitEvent(st, ctx->aux_join[i], 0));
    }
    uint64_t* h_ext;
    ZKW_TRY(sc.get(eb, (void**)&h_ext));
    {
        std::vector<const uint64_t*> e_const(F), e_q(A), e_sig(pk->nperm);
        for (unsigned c = 0; c < F; c++) e_const[c] = pk->fixed_cosets[c];
        for (unsigned c = 0; c < A; c++) e_q[c] = pk->fixed_cosets[pk->q_enable_col(c)];
        for (unsigned c = 0; c < pk->nperm; c++) e_sig[c] = pk->sigma_cosets[c];
        zkw_quotient_inputs qi;
        memset(&qi, 0, sizeof(qi));
        qi.shape = sh;
        qi.advice = e_adv.data(); qi.constants = e_const.data(); qi.table = pk->fixed_cosets[pk->table_col()];
        qi.q_enable = e_q.data(); qi.q_lookup = L ? nullptr : pk->fixed_cosets[pk->q_lookup_col()];
        qi.sigma = e_sig.data(); qi.perm_z = e_pz.data(); qi.lookup_z = e_lz.data(); qi.lookup_a = e_la.data(); qi.lookup_s = e_ls.data();
        qi.l0 = pk->l0_coset; qi.l_last = pk->l_last_coset; qi.l_active = pk->l_active_coset;
        memcpy(qi.y, y.l, 32); memcpy(qi.beta, beta.l, 32); memcpy(qi.gamma, gamma.l, 32); memcpy(qi.theta, theta.l, 32);
        ZKW_TRY(quotient_run(ctx, &qi, h_ext));
        ZKW_TRY(ntt_run(ctx, h_ext, sh.ext_k, h_ext, sh.ext_k, dom.dc.ext_omega_inv, false, dom.dc.ext_scale3));
    }
    const unsigned pieces = sh.cs_degree - 1;
    {
        std::vector<std::pair<int, const uint64_t*>> polys;
        for (unsigned i = 0; i < pieces; i++) polys.push_back({ZKW_BASES_G, h_ext + 4 * (size_t)i * n});
        ZKW_TRY(commit_batch(ctx, tr, polys, n));
    }
    const Fr x = tr.squeeze();

    // ---- 7. evaluations ----
    // distinct rotations, in order of first appearance in the query list (the GWC point sets)
    struct Query { int rot; const uint64_t* poly; };
    std::vector<Query> queries;
    const int last_rot = -((int)sh.blinding_factors + 1);
    for (unsigned c = 0; c < A; c++) for (int r = 0; r < 4; r++) queries.push_back({r, adv_cf[c]});
    for (unsigned l = 0; l < L; l++) queries.push_back({0, adv_cf[A + l]});
    for (unsigned s = 0; s < nsets; s++) { queries.push_back({0, perm_z_cf[s]}); queries.push_back({1, perm_z_cf[s]}); }
    for (int s = (int)nsets - 2; s >= 0; s--) queries.push_back({last_rot, perm_z_cf[s]});
    for (unsigned l = 0; l < nlk; l++) {
        queries.push_back({0, lk_z_cf[l]}); queries.push_back({0, lk_a_cf[l]}); queries.push_back({0, lk_s_cf[l]});
        queries.push_back({-1, lk_a_cf[l]}); queries.push_back({1, lk_z_cf[l]});
    }
    for (unsigned c = 0; c < pk->nfixed; c++) queries.push_back({0, pk->fixed_polys[c]});
    for (unsigned c = 0; c < pk->nperm; c++) queries.push_back({0, pk->sigma_polys[c]});
    // h(X) = sum_i x^(n i) h_i(X)
    uint64_t* h_poly;
    ZKW_TRY(sc.get(vb, (void**)&h_poly));
    const Fr xn = x.pow((uint64_t)n);
    uint64_t *d_ptrs, *d_weights;
    const size_t max_terms = queries.size() + 8;
    ZKW_TRY(sc.get(max_terms * 8, (void**)&d_ptrs));
    ZKW_TRY(sc.get(max_terms * 32, (void**)&d_weights));
    auto lincomb = [&](const std::vector<const uint64_t*>& polys, const std::vector<Fr>& w, uint64_t* outp) -> int {
        ZKW_CUDA(ctx, cudaMemcpyAsync(d_ptrs, polys.data(), polys.size() * 8, cudaMemcpyHostToDevice, st));
        ZKW_CUDA(ctx, cudaMemcpyAsync(d_weights, w.data(), w.size() * 32, cudaMemcpyHostToDevice, st));
        LinCombArgs a;
        a.polys = (const uint4* const*)d_ptrs; a.weights = (const uint4*)d_weights; a.npolys = (int)polys.size(); a.n = n;
        { ProfScope ps_(ctx, "lincomb_kernel"); lincomb_kernel<<<grid_for(n, 128), 128, 0, st>>>(a, (uint4*)outp); }
        ZKW_LAUNCHED(ctx);
        return cudaStreamSynchronize(st) == cudaSuccess ? ZKW_OK : ZKW_ERR_CUDA;  // host vectors may die after return
    };
    {
        std::vector<const uint64_t*> ps;
        std::vector<Fr> ws;
        Fr w = Fr::one();
        for (unsigned i = 0; i < pieces; i++) { ps.push_back(h_ext + 4 * (size_t)i * n); ws.push_back(w); w = w * xn; }
        ZKW_TRY(lincomb(ps, ws, h_poly));
    }
    queries.push_back({0, h_poly});
    queries.push_back({0, random_poly});
    std::vector<int> rots;
    for (auto& q : queries) if (std::find(rots.begin(), rots.end(), q.rot) == rots.end()) rots.push_back(q.rot);

    // batched polynomial evaluation: (poly, point) pairs -> values, two reduction levels on the device
    auto eval_many = [&](const std::vector<std::pair<const uint64_t*, Fr>>& reqs, std::vector<Fr>& vals) -> int {
        Scratch es(ctx);
        std::vector<Fr> pts;
        std::vector<EvalJob> jobs(reqs.size());
        for (size_t i = 0; i < reqs.size(); i++) {
            size_t pi = 0;
            while (pi < pts.size() && pts[pi] != reqs[i].second) pi++;
            if (pi == pts.size()) pts.push_back(reqs[i].second);
            jobs[i].coeffs = (const uint4*)reqs[i].first;
            jobs[i].point = (uint32_t)pi;
        }
        std::vector<Fr> pows(pts.size() * 40);
        for (size_t pi = 0; pi < pts.size(); pi++) {
            Fr w = pts[pi];
            for (int l = 0; l < 40; l++) { pows[pi * 40 + l] = w; w = w.sqr(); }
        }
        uint64_t *d_pows, *d_part0, *d_part1;
        EvalJob* d_jobs;
        const size_t nb0 = (n + kScanBlock - 1) / kScanBlock, nb1 = (nb0 + kScanBlock - 1) / kScanBlock;
        ZKW_TRY(es.get(pows.size() * 32, (void**)&d_pows));
        ZKW_TRY(es.get(jobs.size() * sizeof(EvalJob), (void**)&d_jobs));
        ZKW_TRY(es.get(jobs.size() * nb0 * 32, (void**)&d_part0));
        ZKW_TRY(es.get(jobs.size() * nb1 * 32, (void**)&d_part1));
        ZKW_CUDA(ctx, cudaMemcpyAsync(d_pows, pows.data(), pows.size() * 32, cudaMemcpyHostToDevice, st));
        ZKW_CUDA(ctx, cudaMemcpyAsync(d_jobs, jobs.data(), jobs.size() * sizeof(EvalJob), cudaMemcpyHostToDevice, st));
        { ProfScope ps_(ctx, "eval_reduce_kernel"); eval_reduce_kernel<<<dim3((unsigned)nb0, (unsigned)jobs.size()), kScanThreads, 0, st>>>(d_jobs, nullptr, 0, (const uint4*)d_pows, 0, (uint4*)d_part0, n, nb0); }
        ZKW_LAUNCHED(ctx);
        const uint64_t* result = d_part0;
        size_t stride = nb0, count = nb0;
        int log_stride = 11;
        uint64_t* bufs[2] = {d_part1, d_part0};
        int which = 0;
        while (count > 1) {
            const size_t nb = (count + kScanBlock - 1) / kScanBlock;
            { ProfScope ps_(ctx, "eval_reduce_kernel"); eval_reduce_kernel<<<dim3((unsigned)nb, (unsigned)jobs.size()), kScanThreads, 0, st>>>(d_jobs, (const uint4*)result, stride, (const uint4*)d_pows, log_stride, (uint4*)bufs[which], count, nb); }
            ZKW_LAUNCHED(ctx);
            result = bufs[which];
            stride = nb; count = nb;
            which ^= 1;
            log_stride += 11;
        }
        vals.resize(jobs.size());
        ZKW_CUDA(ctx, cudaMemcpy2DAsync(vals.data(), 32, result, stride * 32, 32, jobs.size(), cudaMemcpyDeviceToHost, st));
        ZKW_CUDA(ctx, cudaStreamSynchronize(st));
        return ZKW_OK;
    };
    // everything the proof carries: advice, fixed, random, sigma, perm z, lookups
    {
        std::vector<std::pair<const uint64_t*, Fr>> ev;
        auto at = [&](const uint64_t* poly, int r) { ev.push_back({poly, rotate(dom, x, r)}); };
        for (unsigned c = 0; c < A; c++) for (int r = 0; r < 4; r++) at(adv_cf[c], r);
        for (unsigned l = 0; l < L; l++) at(adv_cf[A + l], 0);
        for (unsigned c = 0; c < pk->nfixed; c++) at(pk->fixed_polys[c], 0);
        at(random_poly, 0);
        for (unsigned c = 0; c < pk->nperm; c++) at(pk->sigma_polys[c], 0);
        for (unsigned s = 0; s < nsets; s++) {
            at(perm_z_cf[s], 0); at(perm_z_cf[s], 1);
            if (s + 1 != nsets) at(perm_z_cf[s], last_rot);
        }
        for (unsigned l = 0; l < nlk; l++) { at(lk_z_cf[l], 0); at(lk_z_cf[l], 1); at(lk_a_cf[l], 0); at(lk_a_cf[l], -1); at(lk_s_cf[l], 0); }
        std::vector<Fr> vals;
        ZKW_TRY(eval_many(ev, vals));
        for (auto& v : vals) tr.write_scalar(v);
    }

    // kate division of an n-coefficient polynomial by (X - z): q has n-1 coefficients, q[n-1] = 0
    uint64_t *kd_terms, *kd_pre;
    ZKW_TRY(sc.get(vb, (void**)&kd_terms));
    ZKW_TRY(sc.get(vb, (void**)&kd_pre));
    auto kate_div = [&](const uint64_t* poly, const Fr& z, uint64_t* q) -> int {
        const Fr zinv = fp_inv_bingcd(z);
        { ProfScope ps_(ctx, "kate_terms_kernel"); kate_terms_kernel<<<grid_for((n + 15) / 16, 128), 128, 0, st>>>((const uint4*)poly, (uint4*)kd_terms, z, n); }
        ZKW_LAUNCHED(ctx);
        ZKW_TRY((scan_run<false, false>(ctx, kd_terms, kd_pre, n, blocks, nullptr)));
        ZKW_CUDA(ctx, cudaMemsetAsync(q + 4 * (n - 1), 0, 32, st));
        { ProfScope ps_(ctx, "kate_finish_kernel"); kate_finish_kernel<<<grid_for((n + 15) / 16, 128), 128, 0, st>>>((const uint4*)kd_pre, (uint4*)q, zinv, n); }
        ZKW_LAUNCHED(ctx);
        return ZKW_OK;
    };

    if (shplonk) {
        // ---- 8'. multi-open (SHPLONK): halo2_proofs::poly::kzg::multiopen::shplonk::ProverSHPLONK ----
        const Fr ych = tr.squeeze();
        const Fr vch = tr.squeeze();
        auto canon_less = [](const Fr& a, const Fr& b) {
            Fr ca = a.from_mont(), cb = b.from_mont();
            for (int i = 7; i >= 0; i--) if (ca.l[i] != cb.l[i]) return ca.l[i] < cb.l[i];
            return false;
        };
        // commitment -> point set (first-appearance order), then point set -> commitments
        struct Comm { const uint64_t* poly; std::vector<Fr> pts; };
        std::vector<Comm> comms;
        std::vector<Fr> super_pts;
        for (auto& q : queries) {
            const Fr pt = rotate(dom, x, q.rot);
            if (std::find(super_pts.begin(), super_pts.end(), pt) == super_pts.end()) super_pts.push_back(pt);
            auto it = std::find_if(comms.begin(), comms.end(), [&](const Comm& c) { return c.poly == q.poly; });
            if (it == comms.end()) comms.push_back({q.poly, {pt}});
            else if (std::find(it->pts.begin(), it->pts.end(), pt) == it->pts.end()) it->pts.push_back(pt);
        }
        for (auto& c : comms) std::sort(c.pts.begin(), c.pts.end(), canon_less);   // BTreeSet order
        struct RSet { std::vector<Fr> pts; std::vector<const uint64_t*> polys; std::vector<std::vector<Fr>> r_x; };
        std::vector<RSet> rsets;
        for (auto& c : comms) {
            auto it = std::find_if(rsets.begin(), rsets.end(), [&](const RSet& r) { return r.pts == c.pts; });
            if (it == rsets.end()) { rsets.push_back({c.pts, {c.poly}, {}}); }
            else it->polys.push_back(c.poly);
        }
        // evaluations of every commitment at the points of its set, then the low-degree equivalents r(X)
        {
            std::vector<std::pair<const uint64_t*, Fr>> ev;
            for (auto& rs : rsets) for (auto poly : rs.polys) for (auto& pt : rs.pts) ev.push_back({poly, pt});
            std::vector<Fr> vals;
            ZKW_TRY(eval_many(ev, vals));
            size_t k = 0;
            for (auto& rs : rsets) {
                const size_t m = rs.pts.size();
                for (size_t pi = 0; pi < rs.polys.size(); pi++) {
                    // Lagrange interpolation through (pts[i], vals[k + i])
                    std::vector<Fr> out(m, Fr::zero());
                    for (size_t i = 0; i < m; i++) {
                        std::vector<Fr> num(1, Fr::one());
                        Fr den = Fr::one();
                        for (size_t j = 0; j < m; j++) {
                            if (j == i) continue;
                            std::vector<Fr> nn(num.size() + 1, Fr::zero());
                            for (size_t d = 0; d < num.size(); d++) { nn[d + 1] = nn[d + 1] + num[d]; nn[d] = nn[d] - rs.pts[j] * num[d]; }
                            num.swap(nn);
                            den = den * (rs.pts[i] - rs.pts[j]);
                        }
                        const Fr cf = vals[k + i] * fp_inv_bingcd(den);
                        for (size_t d = 0; d < num.size(); d++) out[d] = out[d] + cf * num[d];
                    }
                    rs.r_x.push_back(out);
                    k += m;
                }
            }
        }
        auto host_eval = [](const std::vector<Fr>& c, const Fr& at) { Fr acc = Fr::zero(); for (size_t i = c.size(); i-- > 0;) acc = acc * at + c[i]; return acc; };
        uint64_t *tmp_a, *tmp_b, *h_x, *d_low;
        ZKW_TRY(sc.get(vb, (void**)&tmp_a)); ZKW_TRY(sc.get(vb, (void**)&tmp_b)); ZKW_TRY(sc.get(vb, (void**)&h_x));
        ZKW_TRY(sc.get(8 * 32, (void**)&d_low));
        std::vector<uint64_t*> qsets(rsets.size());
        for (size_t i = 0; i < rsets.size(); i++) {
            RSet& rs = rsets[i];
            // N_i(X) = sum_j y^j (P_ij(X) - r_ij(X))
            std::vector<const uint64_t*> ps(rs.polys.begin(), rs.polys.end());
            std::vector<Fr> ws;
            std::vector<Fr> low(rs.pts.size(), Fr::zero());
            Fr py = Fr::one();
            for (size_t j = 0; j < rs.polys.size(); j++) {
                ws.push_back(py);
                for (size_t d = 0; d < low.size(); d++) low[d] = low[d] + py * rs.r_x[j][d];
                py = py * ych;
            }
            ZKW_TRY(lincomb(ps, ws, tmp_a));
            ZKW_CUDA(ctx, cudaMemcpyAsync(d_low, low.data(), low.size() * 32, cudaMemcpyHostToDevice, st));
            { ProfScope ps_(ctx, "sub_low_kernel"); sub_low_kernel<<<1, 32, 0, st>>>((uint4*)tmp_a, (const uint4*)d_low, (int)low.size()); }
            ZKW_LAUNCHED(ctx);
            ZKW_CUDA(ctx, cudaStreamSynchronize(st));
            // Q_i = N_i / prod (X - p): successive synthetic divisions
            ZKW_TRY(sc.get(vb, (void**)&qsets[i]));
            uint64_t* src = tmp_a;
            for (size_t pi = 0; pi < rs.pts.size(); pi++) {
                uint64_t* dst = (pi + 1 == rs.pts.size()) ? qsets[i] : (src == tmp_a ? tmp_b : tmp_a);
                ZKW_TRY(kate_div(src, rs.pts[pi], dst));
                src = dst;
            }
        }
        {
            std::vector<const uint64_t*> ps(qsets.begin(), qsets.end());
            std::vector<Fr> ws;
            Fr pv = Fr::one();
            for (size_t i = 0; i < rsets.size(); i++) { ws.push_back(pv); pv = pv * vch; }
            ZKW_TRY(lincomb(ps, ws, h_x));
        }
        ZKW_TRY(commit_batch(ctx, tr, {{ZKW_BASES_G, h_x}}, n));
        const Fr uch = tr.squeeze();
        // L(X) = sum_i v^i z_i sum_j y^j (P_ij(X) - r_ij(u)) - Z_T(u) h(X), scaled by 1 / z_0
        std::vector<Fr> zdiff(rsets.size());
        for (size_t i = 0; i < rsets.size(); i++) {
            Fr z = Fr::one();
            for (auto& d : super_pts) if (std::find(rsets[i].pts.begin(), rsets[i].pts.end(), d) == rsets[i].pts.end()) z = z * (uch - d);
            zdiff[i] = z;
        }
        Fr zt = Fr::one();
        for (auto& d : super_pts) zt = zt * (uch - d);
        const Fr inv0 = fp_inv_bingcd(zdiff[0]);
        std::vector<const uint64_t*> ps;
        std::vector<Fr> ws;
        Fr cst = Fr::zero();
        Fr pv = Fr::one();
        for (size_t i = 0; i < rsets.size(); i++) {
            Fr py = Fr::one();
            const Fr wi = pv * zdiff[i] * inv0;
            for (size_t j = 0; j < rsets[i].polys.size(); j++) {
                ps.push_back(rsets[i].polys[j]);
                ws.push_back(wi * py);
                cst = cst + wi * py * host_eval(rsets[i].r_x[j], uch);
                py = py * ych;
            }
            pv = pv * vch;
        }
        ps.push_back(h_x);
        ws.push_back((zt * inv0).neg());
        ZKW_TRY(lincomb(ps, ws, tmp_a));
        ZKW_CUDA(ctx, cudaMemcpyAsync(d_low, cst.l, 32, cudaMemcpyHostToDevice, st));
        { ProfScope ps_(ctx, "sub_low_kernel"); sub_low_kernel<<<1, 32, 0, st>>>((uint4*)tmp_a, (const uint4*)d_low, 1); }
        ZKW_LAUNCHED(ctx);
        ZKW_CUDA(ctx, cudaStreamSynchronize(st));
        ZKW_TRY(kate_div(tmp_a, uch, tmp_b));
        ZKW_TRY(commit_batch(ctx, tr, {{ZKW_BASES_G, tmp_b}}, n));
        *out_len = tr.out.size();
        if (!out || out_cap < tr.out.size()) return ZKW_ERR_INVALID;
        memcpy(out, tr.out.data(), tr.out.size());
        return ZKW_OK;
    }

    // ---- 8. multi-open (GWC) ----
    const Fr v = tr.squeeze();
    {
        uint64_t* batch;
        ZKW_TRY(sc.get(vb, (void**)&batch));
        LanePipe wpipe(ctx, tr, n);
        for (int r : rots) {
            uint64_t* wit;
            ZKW_TRY(sc.get(vb, (void**)&wit));
            std::vector<const uint64_t*> ps;
            std::vector<Fr> ws;
            Fr pv = Fr::one();
            for (auto& q : queries) if (q.rot == r) { ps.push_back(q.poly); ws.push_back(pv); pv = pv * v; }
            ZKW_TRY(lincomb(ps, ws, batch));
            ZKW_TRY(kate_div(batch, rotate(dom, x, r), wit));
            ZKW_TRY(wpipe.submit(ZKW_BASES_G, wit));
        }
        ZKW_TRY(wpipe.flush());
    }
    *out_len = tr.out.size();
    if (!out || out_cap < tr.out.size()) return ZKW_ERR_INVALID;
    memcpy(out, tr.out.data(), tr.out.size());
    return ZKW_OK;
}

}  // extern "C"
