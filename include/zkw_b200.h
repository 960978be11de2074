/*
 * zkw_b200.h — C ABI of the B200-native Halo2 prover hot path (BN254 G1 MSM, BN254 Fr NTT,
 * coset quotient evaluation) for zkwebauthn/webauthn-halo2's P-256 ECDSA circuit.
 *
 * This is the drop-in boundary: the entry points below are what a patched
 * `halo2_proofs` (reached from the reference at
 * halo2-circuits/src/ecc/ecdsa_p256.rs:366-373, :416-423, :555-562 [create_proof] and
 * :259-260 [keygen_vk / keygen_pk]) binds over Rust FFI in place of
 *
 *   halo2_proofs::arithmetic::best_multiexp        -> zkw_msm_bn254_g1
 *   halo2_proofs::arithmetic::best_fft             -> zkw_ntt_bn254_fr
 *   poly::EvaluationDomain::lagrange_to_coeff      -> zkw_lagrange_to_coeff
 *   poly::EvaluationDomain::coeff_to_extended      -> zkw_coeff_to_extended
 *   poly::EvaluationDomain::extended_to_coeff      -> zkw_extended_to_coeff
 *   plonk::evaluation::Evaluator::evaluate_h
 *     + EvaluationDomain::divide_by_vanishing_poly -> zkw_quotient_ecdsa
 *
 * (those crates are un-vendored dependencies of the reference, Cargo.toml:12-15; see
 * INTEGRATION.md for the Rust `extern "C"` block and the three patched call sites).
 *
 * Conventions
 *   - Plain C: pointers + sizes, no C++/torch types.  Every function returns ZKW_OK (0) or a
 *     negative zkw_status; nothing aborts.  zkw_strerror() names a status.
 *   - Field elements are halo2curves' in-memory form: [u64;4] little-endian limbs in
 *     Montgomery form (R = 2^256).  G1Affine = {x: Fq, y: Fq}, 64 contiguous bytes, identity
 *     encoded (0,0).  G1 (projective output) = {x,y,z} Jacobian, 96 bytes, identity z = 0.
 *   - Pointer arguments named `*_host`/unsuffixed are HOST memory; `_dev` entry points take
 *     DEVICE pointers valid on the context's device and enqueue on the context's stream
 *     without synchronising (call zkw_ctx_sync).
 *   - A zkw_ctx is bound to one CUDA device; use one ctx per GPU (one process per GPU in the
 *     multi-GPU layout).  Calls on one ctx must be serialised by the caller.
 *   - There is NO CPU fallback: without a usable CUDA device every call returns
 *     ZKW_ERR_NO_DEVICE.
 */
#ifndef ZKW_B200_H
#define ZKW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct zkw_ctx zkw_ctx;

typedef enum {
    ZKW_OK = 0,
    ZKW_ERR_NO_DEVICE = -1,   /* no CUDA device / driver: the product path has no CPU fallback */
    ZKW_ERR_CUDA = -2,        /* a CUDA runtime call failed; zkw_last_cuda_error() has the text */
    ZKW_ERR_INVALID = -3,     /* bad argument (NULL, size not a power of two, k out of range …) */
    ZKW_ERR_OOM = -4,         /* device allocation failed */
    ZKW_ERR_STATE = -5,       /* e.g. MSM against SRS bases before zkw_srs_load */
    ZKW_ERR_UNSUPPORTED = -6, /* circuit shape outside what the quotient kernel was built for */
    ZKW_ERR_SIGNATURE = -7    /* the assertion's signature does not verify: the ECDSA circuit has no satisfying assignment */
} zkw_status;

/* ---- context ------------------------------------------------------------------------- */
int zkw_ctx_create(int device, zkw_ctx** out);
void zkw_ctx_destroy(zkw_ctx* ctx);
int zkw_ctx_sync(zkw_ctx* ctx);
/* The CUDA stream (cudaStream_t as void*) every `_dev` call of this ctx is enqueued on. */
void* zkw_ctx_stream(zkw_ctx* ctx);
const char* zkw_strerror(int status);
const char* zkw_last_cuda_error(zkw_ctx* ctx);
/* Number of kernel launches this ctx has issued since creation (bench.py's gpu_launches). */
uint64_t zkw_ctx_launch_count(zkw_ctx* ctx);

/* Optional per-kernel device timing (CUDA events on the ctx stream), used by bench.py for the
 * roofline of the dominant kernel.  Kernel names are the __global__ function names
 * ("msm_accumulate_kernel", "ntt_pass_kernel", "quotient_kernel", ...). */
/* Device self-test of the hand-written field arithmetic variants: the dedicated Montgomery squaring against the general
 * product, limb for limb, on `count` pseudo-random values (plus corner cases) of Fr and Fq; mismatches_out = {Fr, Fq} counts. */
int zkw_selftest_field(zkw_ctx* ctx, uint64_t seed, unsigned count, unsigned mismatches_out[2]);

int zkw_profile_enable(zkw_ctx* ctx, int on);
/* Time only launches of one kernel (NULL: every kernel): two events per launch perturb a proof of 215 launches by ~1 ms,
 * bench.py times just the dominant kernel inside its timed region. */
int zkw_profile_filter(zkw_ctx* ctx, const char* kernel_or_null);
int zkw_profile_reset(zkw_ctx* ctx);
int zkw_profile_read(zkw_ctx* ctx, const char* kernel, double* total_ms, uint64_t* launches);
int zkw_profile_names(zkw_ctx* ctx, char* buf, size_t cap); /* comma-separated names seen so far */

/* ---- SRS residency: replaces ParamsKZG{g, g_lagrange} living in host Vec<G1Affine> ------ */
/* Copies n affine points of each basis to the device once; later MSMs name them by id. */
int zkw_srs_load(zkw_ctx* ctx, const uint64_t* g /* n*8 */, const uint64_t* g_lagrange /* n*8 or NULL */, size_t n);

/* Development SRS generated on the device for a caller-chosen tau (Montgomery form): the
 * replacement of halo2-lib's gen_srs(k) / ParamsKZG::setup (ecdsa_p256.rs:258,279,338,388,430):
 * g[i] = tau^i G and g_lagrange[i] = L_i(tau) G, n = 2^k, both resident (with window tables). */
int zkw_srs_setup(zkw_ctx* ctx, unsigned k, const uint64_t tau[4]);
/* Copy the first n points of a resident basis back to the host (n*8 u64). */
int zkw_srs_get(zkw_ctx* ctx, int which_bases, uint64_t* out_xy, size_t n);
/* out[i] = scalars[i] * G (G1 generator), affine; host pointers. */
int zkw_g1_fixed_base_mul(zkw_ctx* ctx, const uint64_t* scalars, size_t n, uint64_t* out_xy);
/* Same, from device arrays (the context keeps its own copies). */
int zkw_srs_load_dev(zkw_ctx* ctx, const uint64_t* g_dev, const uint64_t* g_lagrange_dev, size_t n);
/* MSM tuning, effective for bases loaded afterwards: window_bits 0 = automatic; precompute != 0 keeps
 * 2^(c*w)*P_i for every window of the resident bases (c.f. msm.cu). Environment overrides at ctx
 * creation: ZKW_MSM_WINDOW_BITS, ZKW_MSM_PRECOMPUTE. */
int zkw_msm_config(zkw_ctx* ctx, int window_bits, int precompute);
/* Window width c an n-point MSM uses with the current configuration (the MSM makes ceil(255 / c) mixed additions per
 * point); negative = error.  Reporting aid: bench.py derives its product counts from it. */
int zkw_msm_window_bits(zkw_ctx* ctx, size_t n);

enum { ZKW_BASES_G = 0, ZKW_BASES_G_LAGRANGE = 1, ZKW_BASES_CALLER = 2 };

/* ---- MSM: halo2_proofs::arithmetic::best_multiexp(coeffs, bases) -> C::Curve -------------- */
/* scalars: n*4 u64, Montgomery form as stored by halo2curves (the kernel de-Montgomerises,
 * the analogue of upstream's `to_repr()`).  bases: n*8 u64 when which_bases == ZKW_BASES_CALLER,
 * else ignored.  out_xyz: Jacobian (X,Y,Z) Montgomery, identity => Z = 0.  The representative
 * returned is the normalised one, (x, y, 1), so the bytes are deterministic. */
int zkw_msm_bn254_g1(zkw_ctx* ctx, int which_bases, const uint64_t* bases, const uint64_t* scalars,
                     size_t n, uint64_t out_xyz[12]);
/* Device-pointer variant: scalars_dev (and bases_dev for ZKW_BASES_CALLER) are device pointers;
 * out_xyz_dev receives 12 u64 on the device; asynchronous on the ctx stream. */
int zkw_msm_bn254_g1_dev(zkw_ctx* ctx, int which_bases, const uint64_t* bases_dev,
                         const uint64_t* scalars_dev, size_t n, uint64_t* out_xyz_dev);
/* scalars (and caller bases) on the device, result written to host memory; synchronises the ctx
 * stream.  This is the form the prover pipeline uses: commitments feed the host-side transcript. */
int zkw_msm_bn254_g1_dev_to_host(zkw_ctx* ctx, int which_bases, const uint64_t* bases_dev,
                                 const uint64_t* scalars_dev, size_t n, uint64_t out_xyz[12]);
/* Jacobian -> affine (x,y) Montgomery, (0,0) for the identity: C::Curve::batch_normalize. */
int zkw_g1_batch_normalize(zkw_ctx* ctx, const uint64_t* xyz /* m*12 */, size_t m, uint64_t* out_xy /* m*8 */);

/* Sum of m Jacobian points on the HOST, normalised to (x, y, 1) (identity: z = 0): folds the per-GPU partial results of
 * one MSM split across GPUs (each rank's zkw_msm_* result, all-gathered: 96 bytes per rank). */
int zkw_g1_sum(const uint64_t* xyz /* m*12 */, size_t m, uint64_t out_xyz[12]);

/* ---- NTT: halo2_proofs::arithmetic::best_fft(a, omega, log_n) ------------------------------ */
/* In place, natural order in and out: a[i] <- sum_j a[j] * omega^(i*j).  scale_or_null, if
 * given, multiplies every output (n^-1 for an inverse transform). */
int zkw_ntt_bn254_fr(zkw_ctx* ctx, uint64_t* a /* 2^log_n * 4 */, unsigned log_n, const uint64_t omega[4],
                     const uint64_t* scale_or_null);
int zkw_ntt_bn254_fr_dev(zkw_ctx* ctx, uint64_t* a_dev, unsigned log_n, const uint64_t omega[4],
                         const uint64_t* scale_or_null);

/* ---- EvaluationDomain mirrors (k = log2 rows, ext_k = log2 extended rows) --------------------- */
/* lagrange_to_coeff: inverse NTT over the 2^k domain, scaled by n^-1; in place. */
int zkw_lagrange_to_coeff(zkw_ctx* ctx, uint64_t* a, unsigned k);
int zkw_lagrange_to_coeff_dev(zkw_ctx* ctx, uint64_t* a_dev, unsigned k);
/* coeff_to_lagrange: forward NTT over the 2^k domain; in place. */
int zkw_coeff_to_lagrange(zkw_ctx* ctx, uint64_t* a, unsigned k);
int zkw_coeff_to_lagrange_dev(zkw_ctx* ctx, uint64_t* a_dev, unsigned k);
/* coeff_to_extended: a_i *= zeta^(i mod 3), zero-pad 2^k -> 2^ext_k, forward NTT with the
 * extended root.  coeffs: 2^k*4 u64; out: 2^ext_k*4 u64 (must not alias coeffs). */
int zkw_coeff_to_extended(zkw_ctx* ctx, const uint64_t* coeffs, unsigned k, unsigned ext_k, uint64_t* out);
int zkw_coeff_to_extended_dev(zkw_ctx* ctx, const uint64_t* coeffs_dev, unsigned k, unsigned ext_k, uint64_t* out_dev);
/* extended_to_coeff: inverse NTT over 2^ext_k, scale by 2^-ext_k, a_i *= zeta^-(i mod 3); in
 * place over all 2^ext_k entries (upstream then truncates to n*(degree-1); the caller does). */
int zkw_extended_to_coeff(zkw_ctx* ctx, uint64_t* a, unsigned ext_k);
int zkw_extended_to_coeff_dev(zkw_ctx* ctx, uint64_t* a_dev, unsigned ext_k);

/* ---- quotient: Evaluator::evaluate_h + divide_by_vanishing_poly for the ECDSA circuit ------- */
/* Shape of halo2-lib's FlexGate + Range configuration as instantiated by
 * ECDSACircuit::configure (ecdsa_p256.rs:94-115) from the JSON configs
 * (halo2-circuits/src/configs/ *.config files).  The constraint list is the one the reference's
 * generated verifier checks (proving-server/P256Verifier.yul:406-547):
 *   gates      : for each gate advice column c:  q_c * (a_c(X) + a_c(wX)*a_c(w^2X) - a_c(w^3X))
 *   permutation: columns [constants.., gate advice.., lookup advice..], chunks of cs_degree-2
 *   lookups    : one per lookup advice column: (a_l ; table), or — selector mode, when
 *                num_lookup_advice == 0 — a single lookup (q_lookup * a_0 ; table).           */
typedef struct {
    uint32_t k;                  /* log2 rows                                                 */
    uint32_t ext_k;              /* log2 extended rows (k+2 for every config of the reference)  */
    uint32_t num_advice;         /* gate advice columns A                                     */
    uint32_t num_lookup_advice;  /* dedicated lookup advice columns L; 0 = selector mode       */
    uint32_t num_fixed;          /* constant (fixed, equality-enabled) columns F              */
    uint32_t blinding_factors;   /* cs.blinding_factors(); 6 for this circuit (yul:309-323)      */
    uint32_t cs_degree;          /* cs.degree(): 4 (plain lookup input) or 5 (selector mode)    */
    uint32_t reserved;
} zkw_circuit_shape;

static inline uint32_t zkw_shape_perm_columns(const zkw_circuit_shape* s) {
    return s->num_fixed + s->num_advice + s->num_lookup_advice;
}
static inline uint32_t zkw_shape_perm_sets(const zkw_circuit_shape* s) {
    uint32_t chunk = s->cs_degree - 2;
    return (zkw_shape_perm_columns(s) + chunk - 1) / chunk;
}
static inline uint32_t zkw_shape_lookups(const zkw_circuit_shape* s) {
    return s->num_lookup_advice ? s->num_lookup_advice : 1;
}

/* Every pointer is a 2^ext_k * 4 u64 evaluation vector over the extended coset zeta*<w_ext>
 * (what coeff_to_extended returns), except the four challenges.  Arrays of pointers are indexed
 * as the comments say.  For the `_dev` entry point all vectors are device pointers (the pointer
 * TABLES themselves stay in host memory). */
typedef struct {
    zkw_circuit_shape shape;
    const uint64_t* const* advice;        /* [A + L]  gate advice then lookup advice            */
    const uint64_t* const* constants;     /* [F]                                               */
    const uint64_t* table;                /* range-lookup table column                         */
    const uint64_t* const* q_enable;      /* [A]      gate selectors                           */
    const uint64_t* q_lookup;             /* selector mode only, else NULL                     */
    const uint64_t* const* sigma;         /* [F + A + L] permutation cosets, permutation order */
    const uint64_t* const* perm_z;        /* [perm_sets] grand products                        */
    const uint64_t* const* lookup_z;      /* [lookups]                                         */
    const uint64_t* const* lookup_a;      /* [lookups] permuted input  A'                      */
    const uint64_t* const* lookup_s;      /* [lookups] permuted table  S'                      */
    const uint64_t* l0;
    const uint64_t* l_last;
    const uint64_t* l_active;             /* 1 - (l_last + l_blind)                            */
    uint64_t y[4], beta[4], gamma[4], theta[4];  /* Montgomery form                           */
} zkw_quotient_inputs;

/* h_ext[i] = ( sum_j y^(m-1-j) C_j(row i) ) / ((zeta*w_ext^i)^n - 1), 2^ext_k * 4 u64. */
int zkw_quotient_ecdsa(zkw_ctx* ctx, const zkw_quotient_inputs* in, uint64_t* h_ext);
int zkw_quotient_ecdsa_dev(zkw_ctx* ctx, const zkw_quotient_inputs* in, uint64_t* h_ext_dev);

/* ---- keygen + create_proof on the device ------------------------------
 * The callers of the hot path, widened onto the device: what the reference reaches through
 *   download_keys      -> keygen_vk / keygen_pk                      (ecdsa_p256.rs:256-272)
 *   generate_proof_evm -> create_proof<.., ProverGWC, EvmTranscript> (ecdsa_p256.rs:329-377)
 *   generate_proof     -> create_proof<.., Blake2bWrite>             (ecdsa_p256.rs:379-427)
 * The resident SRS (zkw_srs_load / zkw_srs_setup) must have exactly n = 2^k points. */
typedef struct zkw_pk zkw_pk;

/* fixed_values: [num_fixed constants, table, num_advice gate selectors, (q_lookup in selector mode)],
 * each n*4 u64 Montgomery, host.  perm_mapping: one array per permutation column
 * [constants.., gate advice.., lookup advice..], each n pairs (col', row') of u32 — the cycle
 * successor mapping of halo2's keygen Assembly.  Everything derived (polys, extended cosets, sigma,
 * l_0 / l_last / l_active, sorted lookup table, VK commitments) is computed and kept on the device. */
int zkw_keygen(zkw_ctx* ctx, const zkw_circuit_shape* shape, const uint64_t* const* fixed_values,
               const uint32_t* const* perm_mapping, zkw_pk** out);
void zkw_pk_destroy(zkw_ctx* ctx, zkw_pk* pk);
int zkw_pk_info(const zkw_pk* pk, uint32_t* num_fixed_cols, uint32_t* num_perm_cols);
/* VK: commitments to the fixed and sigma columns (affine Montgomery, 8 u64 each) and the transcript
 * digest (Montgomery Fr).  The digest is self-defined (upstream hashes the Rust Debug string of the
 * pinned VK): Blake2b-512("Halo2-Verify-Key") over a canonical serialisation, mod r. */
int zkw_pk_vk(const zkw_pk* pk, uint64_t* fixed_commitments_xy, uint64_t* perm_commitments_xy, uint64_t digest[4]);

/* Key files: ProvingKey / VerifyingKey ::to_bytes(SerdeFormat::RawBytes) + ::read (ecdsa_p256.rs:261-270 writes them in
 * download_keys, :339-343 and :388-393 read the proving key on every request, :280-284 the verifying key).  Raw bytes
 * = in-memory Montgomery limbs.  The proving-key file starts with the verifying key (shape, digest, commitments),
 * followed by values + polynomial of every fixed and permutation column; extended cosets are recomputed on load (one
 * NTT each).  zkw_pk_read needs the resident SRS of the same size.  zkw_vk_read accepts either file. */
int zkw_pk_write(zkw_ctx* ctx, const zkw_pk* pk, const char* path);
int zkw_pk_read(zkw_ctx* ctx, const char* path, zkw_pk** out);
int zkw_vk_write(const zkw_pk* pk, const char* path);
int zkw_vk_read(const char* path, zkw_circuit_shape* shape, uint32_t* num_fixed_cols, uint32_t* num_perm_cols,
                uint64_t* fixed_commitments_xy, size_t fixed_cap, uint64_t* perm_commitments_xy, size_t perm_cap, uint64_t digest[4]);

enum { ZKW_TRANSCRIPT_BLAKE2B = 0, ZKW_TRANSCRIPT_EVM = 1 };
/* advice: [num_advice + num_lookup_advice] host arrays holding the first advice_rows[c] (<= n - 7)
 * rows of each advice column (Montgomery); unassigned usable rows are zero, the last
 * blinding_factors + 1 rows are blinding.  seed keys the blinding stream (upstream: OsRng).
 * Multi-open is GWC under both transcripts.  out receives the proof bytes (*out_len), ZKW_ERR_INVALID
 * if out_cap is too small (out_len is still set) or if a lookup input is not in the table. */
int zkw_create_proof(zkw_ctx* ctx, const zkw_pk* pk, const uint64_t* const* advice, const size_t* advice_rows,
                     uint64_t seed, int transcript, uint8_t* out, size_t out_cap, size_t* out_len);

/* flags: the advice arrays live in device memory / hold canonical integers (< 2r, little-endian u64
 * limbs) that the device converts to Montgomery form; the multi-open argument is SHPLONK (what the
 * reference's generate_proof uses, ecdsa_p256.rs:416-423) instead of GWC (generate_proof_evm, :366-373). */
enum { ZKW_ADVICE_ON_DEVICE = 1, ZKW_ADVICE_CANONICAL = 2, ZKW_MULTIOPEN_SHPLONK = 4 /* default: GWC */,
       ZKW_ADVICE_U64 = 8 /* advice arrays hold ONE u64 per row (values < 2^64), widened + Montgomerised on the device */ };
int zkw_create_proof_ex(zkw_ctx* ctx, const zkw_pk* pk, const uint64_t* const* advice, const size_t* advice_rows,
                        uint64_t seed, int transcript, unsigned flags, uint8_t* out, size_t out_cap, size_t* out_len);

/* The same with a 256-bit blinding seed (draw it from the OS: upstream uses OsRng, ecdsa_p256.rs:362,412).  Blinding
 * scalars and the random polynomial come from ChaCha20(seed; nonce = column stream, counter = row), 512 bits per
 * scalar reduced modulo r.  The 64-bit `seed` of the two entry points above is this seed padded with zeros -
 * reproducible streams for tests, NOT zero-knowledge. */
int zkw_create_proof_seeded(zkw_ctx* ctx, const zkw_pk* pk, const uint64_t* const* advice, const size_t* advice_rows,
                            const uint8_t seed[32], int transcript, unsigned flags, uint8_t* out, size_t out_cap, size_t* out_len);

/* The same, with the witness still being produced while the proof starts: the witness-independent device work (the
 * vanishing argument's random polynomial and its commitment, one MSM) is launched first, then `ready(user)` is called once
 * on the calling thread and must return ZKW_OK when the advice arrays may be read (or an error, which aborts the proof
 * and is returned).  zkw_prover_prove uses it to hide the host-side witness synthesis behind that MSM. */
typedef int (*zkw_advice_ready_fn)(void* user);
int zkw_create_proof_overlapped(zkw_ctx* ctx, const zkw_pk* pk, const uint64_t* const* advice, const size_t* advice_rows,
                                const uint8_t seed[32], int transcript, unsigned flags, zkw_advice_ready_fn ready, void* user,
                                uint8_t* out, size_t out_cap, size_t* out_len);

/* TEST SHAPE ONLY (the real circuit is zkw_ecdsa_* below): host-side assignment of a PRNG-filled system with the same
 * gate, lookup and copy-constraint kinds and the column counts of a config line - used for sizes the ECDSA circuit cannot
 * fit (k < 11) and for byte-for-byte comparisons with the Python oracle prover; it attests nothing about a signature.
 * Fills cols_out[c] (c < num_advice: 4 * floor((2^k - blinding_factors - 1) / 4) cells; lookup-advice columns:
 * 2^k - blinding_factors - 1 cells) with canonical values < 2^64 keyed by the assertion bytes, and rows_out[c]
 * (may be NULL) with the cell counts.  Pure host code: no device, no context.  Feed the columns to
 * zkw_create_proof_ex with ZKW_ADVICE_U64. */
int zkw_synth_witness(const zkw_circuit_shape* shape, uint32_t lookup_bits, const uint8_t* assertion, size_t assertion_len,
                      uint64_t* const* cols_out, size_t* rows_out);

/* ---- the P-256 ECDSA verification circuit: ECDSACircuit::{configure, synthesize} (ecdsa_p256.rs:94-206) --------
 * The one-line JSON config of the reference (struct CircuitParams, ecdsa_p256.rs:53-63; strategy is always "Simple"). */
typedef struct {
    uint32_t degree, num_advice, num_lookup_advice, num_fixed, lookup_bits, limb_bits, num_limbs;
} zkw_circuit_params;
typedef struct zkw_ecdsa_circuit zkw_ecdsa_circuit;

/* Lays out the circuit once (keygen's `without_witnesses` pass): selectors, copy constraints, constants, lookup
 * cells.  ZKW_ERR_UNSUPPORTED if it does not fit the 2^degree - 7 usable rows of the given columns (the handle is
 * still returned for zkw_ecdsa_circuit_rows).  Pure host code: no device, no context. */
int zkw_ecdsa_circuit_new(const zkw_circuit_params* params, zkw_ecdsa_circuit** out);
void zkw_ecdsa_circuit_free(zkw_ecdsa_circuit* c);
int zkw_ecdsa_circuit_shape(const zkw_ecdsa_circuit* c, zkw_circuit_shape* out);
/* rows_out[num_advice + lookup columns]: cells assigned per advice column; stats_out (may be NULL): gate cells,
 * lookup cells, distinct constants, copy constraints. */
int zkw_ecdsa_circuit_rows(const zkw_ecdsa_circuit* c, size_t* rows_out, uint64_t stats_out[4]);
/* keygen inputs, in the layout zkw_keygen takes: fixed_out[num_fixed + 1 + num_advice (+1)] arrays of 2^degree * 4
 * u64 CANONICAL integers (convert with zkw_fr_to_mont); mapping_out[num_fixed + num_advice + lookup columns] arrays
 * of 2^degree (col', row') u32 pairs — every copy class closed into a cycle in increasing (col, row) order. */
int zkw_ecdsa_circuit_fixed(const zkw_ecdsa_circuit* c, uint64_t* const* fixed_out);
int zkw_ecdsa_circuit_permutation(const zkw_ecdsa_circuit* c, uint32_t* const* mapping_out);
/* Witness synthesis for one assertion: the five 32-byte little-endian canonical encodings generate_proof{,_evm}
 * take (ecdsa_p256.rs:329,379).  advice_out[c]: rows(c) * 4 u64 CANONICAL integers (zkw_ecdsa_circuit_rows gives the
 * sizes; pass the arrays to zkw_create_proof_ex with ZKW_ADVICE_CANONICAL).  *signature_ok (may be NULL) = 1 iff
 * the assignment satisfies the circuit, i.e. the signature verifies; the cells are written either way. */
int zkw_ecdsa_synthesize(const zkw_ecdsa_circuit* c, const uint8_t pubkey_x[32], const uint8_t pubkey_y[32], const uint8_t r[32],
                         const uint8_t s[32], const uint8_t msg_hash[32], uint64_t* const* advice_out, size_t* rows_out,
                         int* signature_ok);

/* ---- the resident prover: download_keys + generate_proof{,_evm} + the server's concurrent requests ------------------
 * zkw_prover_create = download_keys(degree, pk_path, vk_path) (ecdsa_p256.rs:256-272) kept resident: context on
 * `device`, development SRS for `tau` (Montgomery; gen_srs), the ECDSA circuit for `params`, and the proving key —
 * read from pk_path when that file exists (ProvingKey::read, :339-343), else generated and, when pk_path / vk_path are
 * given, written there (:261-270).  One prover proves one assertion at a time (internally locked). */
typedef struct zkw_prover zkw_prover;
int zkw_prover_create(int device, const zkw_circuit_params* params, const uint64_t tau[4], const char* pk_path, const char* vk_path,
                      zkw_prover** out);
void zkw_prover_destroy(zkw_prover* p);
zkw_ctx* zkw_prover_ctx(zkw_prover* p);
const zkw_pk* zkw_prover_pk(zkw_prover* p);
double zkw_prover_last_synthesis_ms(zkw_prover* p);
/* generate_proof_evm (transcript EVM, flags 0: GWC; ecdsa_p256.rs:329-377) / generate_proof (transcript BLAKE2B, flags
 * ZKW_MULTIOPEN_SHPLONK; :379-427) for one assertion = pubkey_x | pubkey_y | r | s | msg_hash, 5 x 32 little-endian bytes.
 * seed32_or_null: NULL draws the blinding seed from the OS.  ZKW_ERR_INVALID for non-canonical encodings (the reference
 * panics on from_bytes().unwrap()), ZKW_ERR_SIGNATURE when the signature does not verify (nothing is proven). */
int zkw_prover_prove(zkw_prover* p, const uint8_t assertion[160], const uint8_t* seed32_or_null, int transcript, unsigned flags,
                     uint8_t* out, size_t out_cap, size_t* out_len);
/* A batch of independent assertions over `nworkers` provers (same or different GPUs), one host thread per prover pulling
 * from a shared queue: witness synthesis of one proof overlaps the device work of the others.  proofs: count slots of
 * proof_stride bytes; proof_lens[i] = 0 and statuses[i] (may be NULL) = the error for assertions that failed.  Returns
 * the first error, ZKW_OK if every proof was made. */
int zkw_prove_batch(zkw_prover* const* workers, size_t nworkers, const uint8_t* assertions, size_t count, const uint8_t* seeds32_or_null,
                    int transcript, unsigned flags, uint8_t* proofs, size_t proof_stride, size_t* proof_lens, int* statuses);

/* Fr vectors between canonical little-endian integers (< 2r accepted) and halo2curves' Montgomery form;
 * host pointers, n elements of 4 u64. */
int zkw_fr_to_mont(zkw_ctx* ctx, const uint64_t* canonical, uint64_t* out, size_t n);
int zkw_fr_from_mont(zkw_ctx* ctx, const uint64_t* mont, uint64_t* out, size_t n);

/* ---- device memory helpers for FFI callers that keep polynomials resident -------------------- */
int zkw_dev_alloc(zkw_ctx* ctx, size_t bytes, void** out_dev);
int zkw_dev_free(zkw_ctx* ctx, void* dev);
/* page-locked host memory: witness columns synthesised into it reach the device with one asynchronous copy */
int zkw_host_alloc(zkw_ctx* ctx, size_t bytes, void** out_host);
int zkw_host_free(zkw_ctx* ctx, void* host);
int zkw_memcpy_h2d(zkw_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int zkw_memcpy_d2h(zkw_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* ZKW_B200_H */
