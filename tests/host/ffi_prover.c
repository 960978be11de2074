/* ffi_prover.c — the reference's request path driven from plain C11 through include/zkw_b200.h, the way
 * proving-server/src/main.rs would bind it over Rust FFI:
 *   zkw_prover_create  = download_keys(degree, pk_path, vk_path)      halo2-circuits/src/ecc/ecdsa_p256.rs:256-272
 *   zkw_prover_prove   = generate_proof_evm / generate_proof          :329-377 / :379-427
 *   zkw_prove_batch    = the server's concurrent requests             proving-server/src/main.rs:49-79
 * usage: ffi_prover <degree> <num_advice> <num_lookup_advice> <num_fixed> <lookup_bits> <limb_bits> <pk_path> <vk_path>
 *                   <assertions.bin (n x 160 bytes)> <proofs_out.bin>
 * Writes, per assertion, u32 status (0 = ok), u32 length, proof bytes — first for single zkw_prover_prove calls (EVM / GWC),
 * then for one zkw_prove_batch over two provers (the second one READS the key file the first wrote).  Built and run by
 * tests/test_gpu_prover.py, which verifies every proof with the oracle. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "zkw_b200.h"

static int die(const char* what, int rc) {
    fprintf(stderr, "%s: %d (%s)\n", what, rc, zkw_strerror(rc));
    return 1;
}

int main(int argc, char** argv) {
    if (argc != 11) return die("usage", ZKW_ERR_INVALID);
    zkw_circuit_params p;
    p.degree = (uint32_t)atoi(argv[1]); p.num_advice = (uint32_t)atoi(argv[2]); p.num_lookup_advice = (uint32_t)atoi(argv[3]);
    p.num_fixed = (uint32_t)atoi(argv[4]); p.lookup_bits = (uint32_t)atoi(argv[5]); p.limb_bits = (uint32_t)atoi(argv[6]); p.num_limbs = 3;
    const char *pk_path = argv[7], *vk_path = argv[8];
    FILE* fin = fopen(argv[9], "rb");
    if (!fin) return die("assertions", ZKW_ERR_INVALID);
    static unsigned char assertions[64 * 160];
    const size_t count = fread(assertions, 160, 64, fin);
    fclose(fin);
    /* development tau in Montgomery form (the caller chooses it; webauthn-halo2_b200/prover.py uses the same value) */
    const uint64_t tau[4] = {0x3d6c6d4b1b3c5a8eULL, 0, 0, 0};
    uint64_t tau_m[4];
    zkw_prover* a = NULL;
    zkw_prover* b = NULL;
    /* tau must be in Montgomery form: convert through the library (needs a context; the prover's will do after creation, so
     * create a throw-away one first) */
    zkw_ctx* tmp = NULL;
    int rc = zkw_ctx_create(0, &tmp);
    if (rc != ZKW_OK) return die("zkw_ctx_create", rc);
    rc = zkw_fr_to_mont(tmp, tau, tau_m, 1);
    zkw_ctx_destroy(tmp);
    if (rc != ZKW_OK) return die("zkw_fr_to_mont", rc);
    rc = zkw_prover_create(0, &p, tau_m, pk_path, vk_path, &a);          /* keygen on the device + both key files */
    if (rc != ZKW_OK) return die("zkw_prover_create (keygen)", rc);
    rc = zkw_prover_create(0, &p, tau_m, pk_path, NULL, &b);             /* reads the proving key written above */
    if (rc != ZKW_OK) return die("zkw_prover_create (read)", rc);
    FILE* fout = fopen(argv[10], "wb");
    if (!fout) return die("output", ZKW_ERR_INVALID);
    static unsigned char proof[1 << 16];
    unsigned char seed[32];
    for (size_t i = 0; i < count; i++) {
        size_t len = 0;
        memset(seed, (int)(i + 1), sizeof seed);                        /* fixed seeds: the Python side checks determinism */
        rc = zkw_prover_prove(i % 2 ? b : a, assertions + 160 * i, seed, ZKW_TRANSCRIPT_EVM, 0, proof, sizeof proof, &len);
        const uint32_t hdr[2] = {(uint32_t)rc, (uint32_t)(rc == ZKW_OK ? len : 0)};
        fwrite(hdr, 4, 2, fout);
        if (rc == ZKW_OK) fwrite(proof, 1, len, fout);
    }
    /* the batch entry: OS-seeded blinding (seeds = NULL), two provers */
    zkw_prover* workers[2] = {a, b};
    unsigned char* out = (unsigned char*)malloc(count * sizeof proof);
    size_t* lens = (size_t*)calloc(count, sizeof(size_t));
    int* status = (int*)calloc(count, sizeof(int));
    rc = zkw_prove_batch(workers, 2, assertions, count, NULL, ZKW_TRANSCRIPT_EVM, 0, out, sizeof proof, lens, status);
    for (size_t i = 0; i < count; i++) {
        const uint32_t hdr[2] = {(uint32_t)status[i], (uint32_t)lens[i]};
        fwrite(hdr, 4, 2, fout);
        fwrite(out + i * sizeof proof, 1, lens[i], fout);
    }
    fclose(fout);
    printf("FFI_PROVER OK count=%zu batch_rc=%d synth_ms=%.2f\n", count, rc, zkw_prover_last_synthesis_ms(a));
    free(out); free(lens); free(status);
    zkw_prover_destroy(b);
    zkw_prover_destroy(a);
    return 0;
}
