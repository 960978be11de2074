/* ffi_caller.c — a plain C11 caller of include/zkw_b200.h, the way a cgo / Rust `cc` / JNI shim would bind it:
 * proves the header is C (not C++), that the library links with nothing but -lzkw_b200, and that the
 * error-returning entry points behave without a GPU (status codes, never abort).  Built and run by
 * tests/test_abi.py. */
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include "zkw_b200.h"

int main(void) {
    int fails = 0;
    if (strcmp(zkw_strerror(ZKW_OK), "ok") != 0) { printf("strerror(ZKW_OK)\n"); fails++; }
    if (zkw_msm_config(NULL, 0, 1) != ZKW_ERR_INVALID) { printf("msm_config(NULL)\n"); fails++; }
    /* host-only entry point: the synthetic witness for the k = 8 shape of a 1-advice config */
    zkw_circuit_shape sh;
    memset(&sh, 0, sizeof sh);
    sh.k = 8; sh.ext_k = 10; sh.num_advice = 1; sh.num_lookup_advice = 0; sh.num_fixed = 1; sh.blinding_factors = 6; sh.cs_degree = 5;
    const size_t u = (1u << sh.k) - (sh.blinding_factors + 1), cells = 4 * (u / 4);
    uint64_t* col = (uint64_t*)calloc(cells, sizeof(uint64_t));
    uint64_t* cols[1] = {col};
    size_t rows[1] = {0};
    const unsigned char assertion[3] = {1, 2, 3};
    if (zkw_synth_witness(&sh, 7, assertion, sizeof assertion, cols, rows) != ZKW_OK || rows[0] != cells) { printf("synth_witness\n"); fails++; }
    for (size_t g = 0; g < cells / 4 && !fails; g++)   /* every gate a + b*c = d, b in the 2^7 table */
        if (col[4 * g] + col[4 * g + 1] * col[4 * g + 2] != col[4 * g + 3] || col[4 * g + 1] >= 128) { printf("gate %zu\n", g); fails++; }
    if (zkw_synth_witness(NULL, 7, assertion, 3, cols, rows) != ZKW_ERR_INVALID) { printf("synth_witness(NULL)\n"); fails++; }
    /* a context needs a device: on a box without one this is an error code, not a crash or a CPU fallback */
    zkw_ctx* ctx = NULL;
    const int rc = zkw_ctx_create(0, &ctx);
    if (rc == ZKW_OK) { zkw_ctx_destroy(ctx); printf("device present\n"); }
    else if (rc != ZKW_ERR_NO_DEVICE) { printf("ctx_create rc=%d\n", rc); fails++; }
    free(col);
    printf(fails ? "FFI_CALLER FAILED\n" : "FFI_CALLER OK\n");
    return fails ? 1 : 0;
}
