// witness.cu — host-side witness synthesis for the shape-identical synthetic ECDSA circuit
// (webauthn-halo2_b200/circuit.py, class SyntheticEcdsaCircuit; the reference's ECDSACircuit::synthesize,
// halo2-circuits/src/ecc/ecdsa_p256.rs:117-206, runs halo2-ecc's un-vendored chips and cannot be restated).
// One pass over the gates, written straight into the caller's (ideally pinned) buffers: this is the host
// work inside the timed end-to-end path, so it has to be a fraction of the ~28 ms the device needs.
// circuit.py carries the same generator in numpy; tests/test_circuit_cpu.py checks the two agree.
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>
#include "../../include/zkw_b200.h"
#include "hash.hpp"

namespace {

inline uint64_t mix64(uint64_t z) {  // splitmix64 finaliser over a counter
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

}  // namespace

extern "C" int zkw_synth_witness(const zkw_circuit_shape* shape, uint32_t lookup_bits, const uint8_t* assertion, size_t assertion_len,
                                 uint64_t* const* cols_out, size_t* rows_out) {
    if (!shape || !cols_out || (!assertion && assertion_len)) return ZKW_ERR_INVALID;
    const uint64_t n = 1ull << shape->k;
    const uint64_t u = n - (shape->blinding_factors + 1);
    const uint32_t A = shape->num_advice, L = shape->num_lookup_advice;
    uint64_t T = 1ull << lookup_bits;
    if (T > u) T = u;
    const uint64_t G = u / 4;
    uint64_t seed;
    {
        zkw::Blake2b h(64);
        h.update("zkw-b200-synth", 14);
        h.update(assertion, assertion_len);
        uint8_t d[64];
        h.peek_digest(d);
        memcpy(&seed, d, 8);
    }
    const bool pow2 = (T & (T - 1)) == 0;
    const uint64_t nconst = G ? ((G + 1) / 2 < 8 ? (G + 1) / 2 : 8) : 0;
    // gates are independent in pairs (an odd gate chains onto the even gate before it), so a column splits into
    // even-aligned chunks; a few threads bring 2^19 rows from ~1 ms to ~0.3 ms
    const unsigned hw = std::thread::hardware_concurrency();
    const unsigned nthreads = G >= (1u << 15) ? (hw >= 8 ? 4u : (hw >= 2 ? 2u : 1u)) : 1u;
    for (uint32_t c = 0; c < A; c++) {
        uint64_t* col = cols_out[c];
        if (!col) return ZKW_ERR_INVALID;
        const uint64_t s = seed ^ (0x9E3779B97F4A7C15ULL * (uint64_t)(c + 1));
        auto fill = [=](uint64_t g0, uint64_t g1) {
            uint64_t prev_d = 0;
            for (uint64_t g = g0; g < g1; g++) {
                const uint64_t r1 = mix64(s + 3 * g), r2 = mix64(s + 3 * g + 1), r3 = mix64(s + 3 * g + 2);
                uint64_t a = (r2 >> 63) ? ((r2 >> 62) & 1) : (r1 >> 2);   // a mix of bits and wide limbs
                const uint64_t b = pow2 ? (r3 & (T - 1)) : (r3 % T);      // the range-checked cell
                const uint64_t cc = (r2 >> 8) & ((1ull << 40) - 1);
                if (g & 1) a = prev_d;                                    // odd gates chain onto the previous output
                else if (c == 0 && (g >> 1) < nconst) a = (g >> 1) + 1;   // copies of the constants column
                const uint64_t d = a + b * cc;
                col[4 * g] = a; col[4 * g + 1] = b; col[4 * g + 2] = cc; col[4 * g + 3] = d;
                prev_d = d;
            }
        };
        if (nthreads <= 1) {
            fill(0, G);
        } else {
            const uint64_t chunk = ((G + nthreads - 1) / nthreads + 1) & ~1ull;   // even, so chunks start on even gates
            std::vector<std::thread> pool;
            for (unsigned t = 1; t < nthreads; t++) {
                const uint64_t g0 = t * chunk, g1 = g0 + chunk < G ? g0 + chunk : G;
                if (g0 < G) pool.emplace_back(fill, g0, g1);
            }
            fill(0, chunk < G ? chunk : G);
            for (auto& th : pool) th.join();
        }
        if (rows_out) rows_out[c] = 4 * G;
    }
    for (uint32_t l = 0; l < L; l++) {
        uint64_t* col = cols_out[A + l];
        if (!col) return ZKW_ERR_INVALID;
        const uint64_t s = seed ^ (0x9E3779B97F4A7C15ULL * (uint64_t)(A + l + 1));
        for (uint64_t j = 0; j < u; j++) {
            const uint64_t r = mix64(s + j);
            col[j] = pow2 ? (r & (T - 1)) : (r % T);
        }
        if (l == 0) for (uint64_t j = 0; j < G && j < u; j += 3) col[j] = cols_out[0][4 * j + 1];
        if (rows_out) rows_out[A + l] = u;
    }
    return ZKW_OK;
}
