// service.cpp — the resident prover behind the reference's request path, as plain C ABI:
//   download_keys(degree, pk_path, vk_path)            halo2-circuits/src/ecc/ecdsa_p256.rs:256-272   -> zkw_prover_create
//   generate_proof / generate_proof_evm(5 x [u8;32], ..)                      :379-427 / :329-377   -> zkw_prover_prove
//   Rocket's worker threads proving one request each    proving-server/src/main.rs:49-79            -> zkw_prove_batch
// The reference re-reads SRS and proving key from disk on every request (ecdsa_p256.rs:338-343); a zkw_prover keeps
// SRS window tables, proving key, circuit layout and page-locked witness staging resident.  zkw_prove_batch runs
// one std::thread per prover: witness synthesis of one assertion overlaps the device work of the others.
// Host code over the library's own C ABI only (no CUDA calls here).
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#include <sys/random.h>
#include "../../include/zkw_b200.h"

struct zkw_prover {
    zkw_ctx* ctx = nullptr;
    zkw_ecdsa_circuit* circuit = nullptr;
    zkw_pk* pk = nullptr;
    zkw_circuit_shape shape;
    std::vector<size_t> rows;
    std::vector<uint64_t*> staging;      // page-locked, rows[c] * 4 u64 each
    std::mutex mu;                       // one proof at a time per prover (a zkw_ctx is single-threaded)
    double last_synth_ms = 0;
};

namespace {

const uint8_t P256_P_LE[32] = {0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0, 0, 0, 0,
                               0, 0, 0, 0, 0, 0, 0, 0, 0x01, 0, 0, 0, 0xFF, 0xFF, 0xFF, 0xFF};
const uint8_t P256_N_LE[32] = {0x51, 0x25, 0x63, 0xFC, 0xC2, 0xCA, 0xB9, 0xF3, 0x84, 0x9E, 0x17, 0xA7, 0xAD, 0xFA, 0xE6, 0xBC,
                               0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0, 0, 0, 0, 0xFF, 0xFF, 0xFF, 0xFF};

bool less_le(const uint8_t a[32], const uint8_t b[32]) {
    for (int i = 31; i >= 0; i--)
        if (a[i] != b[i]) return a[i] < b[i];
    return false;
}

bool file_exists(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    fclose(f);
    return true;
}

bool os_random(uint8_t* out, size_t n) {
    size_t got = 0;
    while (got < n) {
        ssize_t r = getrandom(out + got, n - got, 0);
        if (r <= 0) return false;
        got += (size_t)r;
    }
    return true;
}

}  // namespace

extern "C" int zkw_prover_create(int device, const zkw_circuit_params* params, const uint64_t tau[4], const char* pk_path, const char* vk_path,
                                 zkw_prover** out) {
    if (!params || !tau || !out) return ZKW_ERR_INVALID;
    *out = nullptr;
    zkw_prover* p = new zkw_prover();
    auto fail = [&](int rc) {
        zkw_prover_destroy(p);
        return rc;
    };
    int rc = zkw_ctx_create(device, &p->ctx);
    if (rc != ZKW_OK) return fail(rc);
    rc = zkw_ecdsa_circuit_new(params, &p->circuit);
    if (rc != ZKW_OK) return fail(rc);
    zkw_ecdsa_circuit_shape(p->circuit, &p->shape);
    const zkw_circuit_shape& sh = p->shape;
    const size_t n = (size_t)1 << sh.k;
    rc = zkw_srs_setup(p->ctx, sh.k, tau);                         // gen_srs(degree)
    if (rc != ZKW_OK) return fail(rc);
    if (pk_path && file_exists(pk_path)) {
        rc = zkw_pk_read(p->ctx, pk_path, &p->pk);                 // ProvingKey::read (ecdsa_p256.rs:339-343)
        if (rc != ZKW_OK) return fail(rc);
    } else {
        const unsigned nfixed = sh.num_fixed + 1 + sh.num_advice + (sh.num_lookup_advice == 0 ? 1 : 0);
        const unsigned nperm = zkw_shape_perm_columns(&sh);
        std::vector<std::vector<uint64_t>> fixed(nfixed, std::vector<uint64_t>(4 * n));
        std::vector<std::vector<uint32_t>> maps(nperm, std::vector<uint32_t>(2 * n));
        std::vector<uint64_t*> fp(nfixed);
        std::vector<uint32_t*> mp(nperm);
        for (unsigned i = 0; i < nfixed; i++) fp[i] = fixed[i].data();
        for (unsigned i = 0; i < nperm; i++) mp[i] = maps[i].data();
        rc = zkw_ecdsa_circuit_fixed(p->circuit, fp.data());
        if (rc == ZKW_OK) rc = zkw_ecdsa_circuit_permutation(p->circuit, mp.data());
        for (unsigned i = 0; i < nfixed && rc == ZKW_OK; i++) rc = zkw_fr_to_mont(p->ctx, fp[i], fp[i], n);
        if (rc == ZKW_OK) rc = zkw_keygen(p->ctx, &sh, fp.data(), mp.data(), &p->pk);   // keygen_vk + keygen_pk (:259-260)
        if (rc != ZKW_OK) return fail(rc);
        if (pk_path) rc = zkw_pk_write(p->ctx, p->pk, pk_path);                          // :261-265
        if (rc != ZKW_OK) return fail(rc);
    }
    if (vk_path) {
        rc = zkw_vk_write(p->pk, vk_path);                                               // :266-270
        if (rc != ZKW_OK) return fail(rc);
    }
    const unsigned ncols = sh.num_advice + sh.num_lookup_advice;
    p->rows.resize(ncols);
    zkw_ecdsa_circuit_rows(p->circuit, p->rows.data(), nullptr);
    p->staging.assign(ncols, nullptr);
    for (unsigned c = 0; c < ncols; c++) {
        void* h = nullptr;
        rc = zkw_host_alloc(p->ctx, 32 * (p->rows[c] ? p->rows[c] : 1), &h);
        if (rc != ZKW_OK) return fail(rc);
        memset(h, 0, 32 * (p->rows[c] ? p->rows[c] : 1));
        p->staging[c] = (uint64_t*)h;
    }
    *out = p;
    return ZKW_OK;
}

extern "C" void zkw_prover_destroy(zkw_prover* p) {
    if (!p) return;
    if (p->ctx) {
        for (uint64_t* h : p->staging)
            if (h) zkw_host_free(p->ctx, h);
        if (p->pk) zkw_pk_destroy(p->ctx, p->pk);
    }
    if (p->circuit) zkw_ecdsa_circuit_free(p->circuit);
    if (p->ctx) zkw_ctx_destroy(p->ctx);
    delete p;
}

extern "C" zkw_ctx* zkw_prover_ctx(zkw_prover* p) { return p ? p->ctx : nullptr; }
extern "C" const zkw_pk* zkw_prover_pk(zkw_prover* p) { return p ? p->pk : nullptr; }
extern "C" double zkw_prover_last_synthesis_ms(zkw_prover* p) { return p ? p->last_synth_ms : 0.0; }

extern "C" int zkw_prover_prove(zkw_prover* p, const uint8_t assertion[160], const uint8_t* seed32_or_null, int transcript, unsigned flags,
                                uint8_t* out, size_t out_cap, size_t* out_len) {
    if (!p || !assertion || !out_len) return ZKW_ERR_INVALID;
    const uint8_t *x = assertion, *y = assertion + 32, *r = assertion + 64, *s = assertion + 96, *m = assertion + 128;
    // Fp::from_bytes / Fq::from_bytes .unwrap() (ecdsa_p256.rs:346-352): canonical encodings only
    if (!less_le(x, P256_P_LE) || !less_le(y, P256_P_LE) || !less_le(r, P256_N_LE) || !less_le(s, P256_N_LE) || !less_le(m, P256_N_LE))
        return ZKW_ERR_INVALID;
    uint8_t seed[32];
    if (seed32_or_null) memcpy(seed, seed32_or_null, 32);
    else if (!os_random(seed, 32)) return ZKW_ERR_STATE;            // OsRng (ecdsa_p256.rs:362,412)
    std::lock_guard<std::mutex> lock(p->mu);
    // the witness is synthesised on a host thread (which fans out to ZKW_SYNTH_THREADS more) while the device already works
    // on the proof's witness-independent part; the proof picks the columns up when `ready` returns
    struct Synth {
        zkw_prover* p;
        const uint8_t *x, *y, *r, *s, *m;
        int rc = ZKW_OK, ok = 0;
        std::thread th;
        static int ready(void* u) {
            Synth* self = (Synth*)u;
            if (self->th.joinable()) self->th.join();
            if (self->rc != ZKW_OK) return self->rc;
            return self->ok ? ZKW_OK : ZKW_ERR_SIGNATURE;            // no satisfying assignment exists: prove nothing
        }
    } sy{p, x, y, r, s, m};
    sy.th = std::thread([&sy] {
        timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        sy.rc = zkw_ecdsa_synthesize(sy.p->circuit, sy.x, sy.y, sy.r, sy.s, sy.m, sy.p->staging.data(), nullptr, &sy.ok);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        sy.p->last_synth_ms = 1e3 * (double)(t1.tv_sec - t0.tv_sec) + 1e-6 * (double)(t1.tv_nsec - t0.tv_nsec);
    });
    flags &= ~(unsigned)(ZKW_ADVICE_ON_DEVICE | ZKW_ADVICE_U64);
    int rc = zkw_create_proof_overlapped(p->ctx, p->pk, p->staging.data(), p->rows.data(), seed, transcript, flags | ZKW_ADVICE_CANONICAL,
                                         &Synth::ready, &sy, out, out_cap, out_len);
    if (sy.th.joinable()) sy.th.join();                              // an early error never reached `ready`
    return rc;
}

extern "C" int zkw_prove_batch(zkw_prover* const* workers, size_t nworkers, const uint8_t* assertions, size_t count, const uint8_t* seeds32_or_null,
                               int transcript, unsigned flags, uint8_t* proofs, size_t proof_stride, size_t* proof_lens, int* statuses) {
    if (!workers || !nworkers || (!assertions && count) || !proofs || !proof_lens) return ZKW_ERR_INVALID;
    for (size_t w = 0; w < nworkers; w++)
        if (!workers[w]) return ZKW_ERR_INVALID;
    std::atomic<size_t> next{0};
    std::atomic<int> first_error{ZKW_OK};
    auto work = [&](zkw_prover* p) {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= count) return;
            size_t len = 0;
            int rc = zkw_prover_prove(p, assertions + 160 * i, seeds32_or_null ? seeds32_or_null + 32 * i : nullptr, transcript, flags,
                                      proofs + proof_stride * i, proof_stride, &len);
            proof_lens[i] = rc == ZKW_OK ? len : 0;
            if (statuses) statuses[i] = rc;
            if (rc != ZKW_OK) {
                int expected = ZKW_OK;
                first_error.compare_exchange_strong(expected, rc);
            }
        }
    };
    std::vector<std::thread> threads;
    for (size_t w = 1; w < nworkers; w++) threads.emplace_back(work, workers[w]);
    work(workers[0]);
    for (auto& t : threads) t.join();
    return first_error.load();
}
