// hash.hpp — host-side Blake2b-512 and Keccak-256 for the two transcripts the reference proves under:
// halo2_proofs' Blake2bWrite<_, _, Challenge255<_>> (generate_proof, halo2-circuits/src/ecc/ecdsa_p256.rs:415)
// and snark-verifier's EvmTranscript (generate_proof_evm, ecdsa_p256.rs:365; framing pinned by
// proving-server/P256Verifier.yul:34-283).  Byte-oriented host code; the hashes cover ~1-3 KB per proof.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace zkw {

// ---- Blake2b (RFC 7693) with personalisation -------------------------------------------------------
class Blake2b {
  public:
    explicit Blake2b(size_t outlen = 64, const char* personal16 = nullptr) : outlen_(outlen) {
        static const uint64_t iv[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                       0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
        for (int i = 0; i < 8; i++) h_[i] = iv[i];
        h_[0] ^= 0x01010000ULL ^ (uint64_t)outlen;
        if (personal16) {
            uint64_t p0, p1;
            memcpy(&p0, personal16, 8);
            memcpy(&p1, personal16 + 8, 8);
            h_[6] ^= p0;
            h_[7] ^= p1;
        }
    }
    void update(const void* data, size_t len) {
        const uint8_t* in = (const uint8_t*)data;
        while (len) {
            if (buflen_ == 128) {
                t_ += 128;
                compress(false);
                buflen_ = 0;
            }
            size_t take = 128 - buflen_ < len ? 128 - buflen_ : len;
            memcpy(buf_ + buflen_, in, take);
            buflen_ += take;
            in += take;
            len -= take;
        }
    }
    void update_byte(uint8_t b) { update(&b, 1); }
    // digest of the data so far without disturbing the running state (Blake2bWrite clones its state)
    void peek_digest(uint8_t* out) const {
        Blake2b c = *this;
        c.t_ += c.buflen_;
        memset(c.buf_ + c.buflen_, 0, 128 - c.buflen_);
        c.compress(true);
        memcpy(out, c.h_, c.outlen_);
    }

  private:
    static uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
    void compress(bool last) {
        static const uint8_t sigma[12][16] = {
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
            {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
            {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
            {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
            {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
        static const uint64_t iv[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                       0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
        uint64_t m[16], v[16];
        memcpy(m, buf_, 128);
        for (int i = 0; i < 8; i++) { v[i] = h_[i]; v[i + 8] = iv[i]; }
        v[12] ^= t_;
        if (last) v[14] = ~v[14];
        auto G = [&](int a, int b, int c, int d, uint64_t x, uint64_t y) {
            v[a] = v[a] + v[b] + x; v[d] = rotr(v[d] ^ v[a], 32);
            v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 24);
            v[a] = v[a] + v[b] + y; v[d] = rotr(v[d] ^ v[a], 16);
            v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 63);
        };
        for (int r = 0; r < 12; r++) {
            const uint8_t* s = sigma[r];
            G(0, 4, 8, 12, m[s[0]], m[s[1]]);   G(1, 5, 9, 13, m[s[2]], m[s[3]]);
            G(2, 6, 10, 14, m[s[4]], m[s[5]]);  G(3, 7, 11, 15, m[s[6]], m[s[7]]);
            G(0, 5, 10, 15, m[s[8]], m[s[9]]);  G(1, 6, 11, 12, m[s[10]], m[s[11]]);
            G(2, 7, 8, 13, m[s[12]], m[s[13]]); G(3, 4, 9, 14, m[s[14]], m[s[15]]);
        }
        for (int i = 0; i < 8; i++) h_[i] ^= v[i] ^ v[i + 8];
    }
    uint64_t h_[8];
    uint64_t t_ = 0;
    uint8_t buf_[128] = {0};
    size_t buflen_ = 0;
    size_t outlen_;
};

// ---- Keccak-256 (original padding 0x01, as the EVM's KECCAK256) ---------------------------------------
inline void keccak_f1600(uint64_t st[25]) {
    static const uint64_t rc[24] = {0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
                                    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
                                    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
                                    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
                                    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
                                    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    static const int rotc[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
    static const int piln[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
    auto rotl = [](uint64_t x, int n) { return (x << n) | (x >> (64 - n)); };
    for (int round = 0; round < 24; round++) {
        uint64_t bc[5];
        for (int i = 0; i < 5; i++) bc[i] = st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20];
        for (int i = 0; i < 5; i++) {
            uint64_t t = bc[(i + 4) % 5] ^ rotl(bc[(i + 1) % 5], 1);
            for (int j = 0; j < 25; j += 5) st[j + i] ^= t;
        }
        uint64_t t = st[1];
        for (int i = 0; i < 24; i++) {
            int j = piln[i];
            uint64_t b = st[j];
            st[j] = rotl(t, rotc[i]);
            t = b;
        }
        for (int j = 0; j < 25; j += 5) {
            for (int i = 0; i < 5; i++) bc[i] = st[j + i];
            for (int i = 0; i < 5; i++) st[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
        }
        st[0] ^= rc[round];
    }
}

inline void keccak256(const uint8_t* data, size_t len, uint8_t out[32]) {
    const size_t rate = 136;
    uint64_t st[25] = {0};
    std::vector<uint8_t> p(data, data + len);
    p.push_back(0x01);
    while (p.size() % rate) p.push_back(0);
    p.back() |= 0x80;
    for (size_t off = 0; off < p.size(); off += rate) {
        for (size_t i = 0; i < rate / 8; i++) {
            uint64_t w;
            memcpy(&w, &p[off + 8 * i], 8);
            st[i] ^= w;
        }
        keccak_f1600(st);
    }
    memcpy(out, st, 32);
}

}  // namespace zkw
