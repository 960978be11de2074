// ecdsa_circuit.cpp — host-side synthesis of the P-256 ECDSA verification circuit: the replacement of
// ECDSACircuit::synthesize (halo2-circuits/src/ecc/ecdsa_p256.rs:117-206, which runs halo2-ecc's
// `ecdsa_verify_no_pubkey_check` with 4-bit windows, :182-191, over halo2-lib's FlexGate / Range / CRT chips)
// in the constraint system the reference's verifier fixes (proving-server/P256Verifier.yul:406-547):
//
//     gate    q_c * (a_c(X) + a_c(wX) * a_c(w^2 X) - a_c(w^3 X))
//     lookup  lookup advice (or q_lookup * a_0 with a single gate column)  in  [0, 2^lookup_bits)
//     copies  permutation over [constants.., gate advice.., lookup advice..]
//
// halo2-lib / halo2-ecc are un-vendored (halo2-circuits/Cargo.toml:12-13), so the cell LAYOUT is this
// repo's own; chip parameters (limb_bits, num_limbs = 3, lookup_bits, column counts) are the reference's JSON
// configs.  Computation:  u1 = m/s, u2 = r/s (mod n);  R = u1*G + u2*PK;  R.x == r;  1 <= r, s < n;  u1, u2 < n;
// PK on the curve.  Unlike the reference (which never asserts the result bit) the result is constrained.
//
// Every relation  sum X*Y - sum X'*Y' + k = 0 (mod m)  is proven as the integer identity
// lhs + 2^258 m = q' m: limb by limb modulo 2^(3 limb_bits) with offset carries, all-positive left / right
// accumulation chains whose ends are copy-constrained, and natively modulo the BN254 scalar field (CRT).
// Point accumulators carry hash-derived constant offset points, so no identity handling is needed; the offsets
// cancel in the final (strict) addition.
//
// Two passes share one code path: the structure pass (zkw_ecdsa_circuit_new; keygen's `without_witnesses`)
// records selectors, copy constraints, constants and lookup cells; the witness pass (zkw_ecdsa_synthesize) only
// writes advice values into the caller's (page-locked) columns.  Pure host code, plain 64-bit integer
// arithmetic; tests compare every cell, selector and permutation entry with oracle/ecdsa_circuit.py.
#include <cstdint>
#include <cstring>
#include <ctime>
#include <chrono>
#include <cstdio>
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <thread>
#include <unordered_map>
#include <vector>
#include "../../include/zkw_b200.h"
#if defined(__x86_64__)
#include <cpuid.h>
#include <immintrin.h>
#endif

namespace {

typedef unsigned __int128 u128;

struct U256 {
    uint64_t l[4];
    bool operator==(const U256& o) const { return l[0] == o.l[0] && l[1] == o.l[1] && l[2] == o.l[2] && l[3] == o.l[3]; }
};
struct U256Hash {
    size_t operator()(const U256& a) const {
        uint64_t h = a.l[0] * 0x9E3779B97F4A7C15ULL;
        h ^= (a.l[1] + 0x7F4A7C15ULL) * 0xBF58476D1CE4E5B9ULL;
        h ^= (a.l[2] + 0x1CE4E5B9ULL) * 0x94D049BB133111EBULL;
        h ^= (a.l[3] + 0x133111EBULL) * 0xD6E8FEB86659FD93ULL;
        return (size_t)(h ^ (h >> 29));
    }
};

inline U256 u256(uint64_t a, uint64_t b = 0, uint64_t c = 0, uint64_t d = 0) { return U256{{a, b, c, d}}; }
inline U256 from128(u128 v) { return u256((uint64_t)v, (uint64_t)(v >> 64)); }
inline bool is_zero(const U256& a) { return !(a.l[0] | a.l[1] | a.l[2] | a.l[3]); }
inline int cmp(const U256& a, const U256& b) {
    for (int i = 3; i >= 0; i--)
        if (a.l[i] != b.l[i]) return a.l[i] < b.l[i] ? -1 : 1;
    return 0;
}
inline uint64_t add_to(U256& a, const U256& b) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a.l[i] + b.l[i]; a.l[i] = (uint64_t)c; c >>= 64; }
    return (uint64_t)c;
}
inline uint64_t sub_from(U256& a, const U256& b) {
    uint64_t borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a.l[i] - b.l[i] - borrow;
        a.l[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
    }
    return borrow;
}
inline U256 shr(const U256& a, unsigned s) {
    U256 r = u256(0);
    unsigned w = s / 64, b = s % 64;
    for (unsigned i = 0; i + w < 4; i++) {
        r.l[i] = a.l[i + w] >> b;
        if (b && i + w + 1 < 4) r.l[i] |= a.l[i + w + 1] << (64 - b);
    }
    return r;
}
inline U256 shl(const U256& a, unsigned s) {
    U256 r = u256(0);
    unsigned w = s / 64, b = s % 64;
    for (int i = 3; i >= (int)w; i--) {
        r.l[i] = a.l[i - w] << b;
        if (b && i - (int)w - 1 >= 0) r.l[i] |= a.l[i - w - 1] >> (64 - b);
    }
    return r;
}
inline U256 pow2(unsigned s) { return shl(u256(1), s); }
inline U256 mask_bits(const U256& a, unsigned bits) {
    if (bits >= 256) return a;
    U256 r = a;
    for (unsigned i = 0; i < 4; i++) {
        if (64 * i >= bits) r.l[i] = 0;
        else if (64 * (i + 1) > bits) r.l[i] &= (~0ULL) >> (64 * (i + 1) - bits);
    }
    return r;
}
inline u128 low128(const U256& a) { return (u128)a.l[0] | ((u128)a.l[1] << 64); }
inline U256 mul128(u128 a, u128 b) {   // 128 x 128 -> 256
    uint64_t a0 = (uint64_t)a, a1 = (uint64_t)(a >> 64), b0 = (uint64_t)b, b1 = (uint64_t)(b >> 64);
    u128 p00 = (u128)a0 * b0, p01 = (u128)a0 * b1, p10 = (u128)a1 * b0, p11 = (u128)a1 * b1;
    U256 r;
    r.l[0] = (uint64_t)p00;
    u128 mid = (p00 >> 64) + (uint64_t)p01 + (uint64_t)p10;
    r.l[1] = (uint64_t)mid;
    u128 hi = (mid >> 64) + (p01 >> 64) + (p10 >> 64) + (uint64_t)p11;
    r.l[2] = (uint64_t)hi;
    r.l[3] = (uint64_t)((hi >> 64) + (p11 >> 64));
    return r;
}
// out[0..7] = a * b
inline void mul_wide(const U256& a, const U256& b, uint64_t out[8]) {
    memset(out, 0, 64);
    for (int i = 0; i < 4; i++) {
        if (!a.l[i]) continue;
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)a.l[i] * b.l[j] + out[i + j]; out[i + j] = (uint64_t)c; c >>= 64; }
        out[i + 4] = (uint64_t)c;
    }
}

// Montgomery arithmetic for an arbitrary odd modulus below 2^256 (P-256's p and n have the top bit set).
struct Mont {
    U256 m, r2, one;
    uint64_t inv;
    void init(const U256& mod) {
        m = mod;
        uint64_t x = 1;
        for (int i = 0; i < 6; i++) x *= 2 - m.l[0] * x;
        inv = 0 - x;
        U256 t = u256(1);
        for (int i = 0; i < 512; i++) {      // 2^512 mod m by doubling
            uint64_t top = t.l[3] >> 63;
            t = shl(t, 1);
            if (top || cmp(t, m) >= 0) sub_from(t, m);
            if (i == 255) one = t;
        }
        r2 = t;
    }
    U256 mul(const U256& a, const U256& b) const {
        uint64_t t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; i++) {
            u128 c = 0;
            for (int j = 0; j < 4; j++) { c += (u128)a.l[j] * b.l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
            c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
            const uint64_t q = t[0] * inv;
            c = (u128)q * m.l[0] + t[0]; c >>= 64;
            for (int j = 1; j < 4; j++) { c += (u128)q * m.l[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
            c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
        }
        U256 r = u256(t[0], t[1], t[2], t[3]);
        if (t[4] || cmp(r, m) >= 0) sub_from(r, m);
        return r;
    }
    U256 to(const U256& a) const { return mul(a, r2); }           // a < m
    U256 from(const U256& a) const { return mul(a, u256(1)); }
    U256 add(const U256& a, const U256& b) const {
        U256 r = a;
        uint64_t c = add_to(r, b);
        if (c || cmp(r, m) >= 0) sub_from(r, m);
        return r;
    }
    U256 sub(const U256& a, const U256& b) const {
        U256 r = a;
        if (sub_from(r, b)) add_to(r, m);
        return r;
    }
    U256 pow(const U256& a, const U256& e) const {   // Montgomery in / out
        U256 acc = one;
        for (int i = 255; i >= 0; i--) {
            acc = mul(acc, acc);
            if ((e.l[i / 64] >> (i % 64)) & 1) acc = mul(acc, a);
        }
        return acc;
    }
    U256 inverse(const U256& a) const {               // Montgomery in / out; 0 -> 0
        U256 e = m;
        sub_from(e, u256(2));
        return pow(a, e);
    }
    U256 reduce(const U256& a) const {                // a < 2^256 -> a mod m
        U256 r = a;
        while (cmp(r, m) >= 0) sub_from(r, m);
        return r;
    }
    U256 mulmod(const U256& a, const U256& b) const { return mul(mul(a, b), r2); }   // canonical in / out
};

const U256 FR_MOD = u256(0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL);
const U256 P256_P = u256(0xFFFFFFFFFFFFFFFFULL, 0x00000000FFFFFFFFULL, 0x0000000000000000ULL, 0xFFFFFFFF00000001ULL);
const U256 P256_N = u256(0xF3B9CAC2FC632551ULL, 0xBCE6FAADA7179E84ULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFF00000000ULL);
const U256 P256_B = u256(0x3BCE3C3E27D2604BULL, 0x651D06B0CC53B0F6ULL, 0xB3EBBD55769886BCULL, 0x5AC635D8AA3A93E7ULL);
const U256 P256_GX = u256(0xF4A13945D898C296ULL, 0x77037D812DEB33A0ULL, 0xF8BCE6E563A440F2ULL, 0x6B17D1F2E12C4247ULL);
const U256 P256_GY = u256(0xCBB6406837BF51F5ULL, 0x2BCE33576B315ECEULL, 0x8EE7EB4A7C0F9E16ULL, 0x4FE342E2FE1A7F9BULL);
// Offset points with unknown discrete logarithm: x = SHA-256(tag || be32(counter)) mod p for the first counter on
// the curve, even y.  Tags "zkw-b200 ecdsa variable-base offset" / "zkw-b200 ecdsa fixed-base offset"; the oracle
// (oracle/ecdsa_circuit.py, offset_point) derives them with hashlib and the parity test compares.
const U256 OFF_VAR_X = u256(0xa356a859026c4881ULL, 0x578d7912c03dd161ULL, 0x52850351ea8e93beULL, 0x64c31fc3e7f7911eULL);
const U256 OFF_VAR_Y = u256(0x1a25aca66e8325e6ULL, 0x3e38168a6420c459ULL, 0xdc1535f0f19193acULL, 0xb4d71455e78ea867ULL);
const U256 OFF_FIX_X = u256(0xa22c3ba538e4bac1ULL, 0x8fa34c0c5184a724ULL, 0x6c294cc857b7717aULL, 0xa9a177813825e976ULL);
const U256 OFF_FIX_Y = u256(0xe64de7c19a3f7010ULL, 0x208acbc2c87bf7edULL, 0xcf2f9549c5b1c89eULL, 0xf991ffe6ed23a1afULL);

constexpr unsigned WINDOW = 4;           // ecdsa_p256.rs:189-190
constexpr unsigned Q_OFFSET_BITS = 258;  // quotients in (-2^258, 2^258), witnessed as q + 2^258

struct Moduli {
    Mont fr, fp, fn;
    Moduli() { fr.init(FR_MOD); fp.init(P256_P); fn.init(P256_N); }
};
const Moduli& moduli() {
    static Moduli m;
    return m;
}

// ---- affine P-256 points in Montgomery form (host-side witness arithmetic) --------------------------------
struct Pt { U256 x, y; };   // Montgomery, never the identity
Pt pt_add(const Mont& f, const Pt& a, const Pt& b) {
    U256 lam = f.mul(f.sub(b.y, a.y), f.inverse(f.sub(b.x, a.x)));
    U256 x = f.sub(f.sub(f.mul(lam, lam), a.x), b.x);
    return Pt{x, f.sub(f.mul(lam, f.sub(a.x, x)), a.y)};
}
Pt pt_dbl(const Mont& f, const Pt& a) {
    U256 xx = f.mul(a.x, a.x);
    U256 three = f.to(u256(3));
    U256 num = f.sub(f.mul(three, xx), three);
    U256 lam = f.mul(num, f.inverse(f.add(a.y, a.y)));
    U256 x = f.sub(f.sub(f.mul(lam, lam), a.x), a.x);
    return Pt{x, f.sub(f.mul(lam, f.sub(a.x, x)), a.y)};
}

// per limb_bits: constant tables of the fixed-base multiplication (canonical coordinates)
struct FixedTables {
    unsigned limb_bits = 0, windows = 0;
    std::vector<U256> x, y;   // [w * 16 + j]
    U256 start_x, start_y;    // -(sum_w 16^w) * OFFSET_VAR
    void build(unsigned LB) {
        const Mont& f = moduli().fp;
        limb_bits = LB;
        windows = (3 * LB + WINDOW - 1) / WINDOW;
        x.resize(16 * windows);
        y.resize(16 * windows);
        Pt base{f.to(P256_GX), f.to(P256_GY)};
        Pt b3{f.to(OFF_FIX_X), f.to(OFF_FIX_Y)};
        Pt off = b3, sum = b3;        // off = 2^w B3; sum = (2^w - 1) B3 once w > 0
        for (unsigned w = 0; w < windows; w++) {
            Pt row = off;
            if (w + 1 == windows) {       // -(2^(W-1) - 1) B3: the offsets of all windows sum to zero
                row = sum;
                row.y = f.sub(u256(0), row.y);
            }
            for (unsigned j = 0; j < 16; j++) {
                if (j) row = pt_add(f, row, base);
                x[16 * w + j] = f.from(row.x);
                y[16 * w + j] = f.from(row.y);
            }
            for (unsigned i = 0; i < WINDOW; i++) base = pt_dbl(f, base);
            if (w + 1 < windows) {
                if (w > 0) sum = pt_add(f, sum, off);
                off = pt_dbl(f, off);
            }
        }
        Pt t{f.to(OFF_VAR_X), f.to(OFF_VAR_Y)};
        Pt vsum = t;
        for (unsigned w = 1; w < windows; w++) {
            for (unsigned i = 0; i < WINDOW; i++) t = pt_dbl(f, t);
            vsum = pt_add(f, vsum, t);
        }
        start_x = f.from(vsum.x);
        start_y = f.from(f.sub(u256(0), vsum.y));
    }
};

struct Elem {
    uint32_t limbs[3];
    uint32_t native;
    U256 value;     // canonical integer < 2^256
    u128 lv[3];
    U256 nat;       // value of the native cell (value mod r)
};
struct Term {
    const Elem* X;
    const Elem* Y;      // nullptr: scalar
    uint64_t scalar;
};
inline Term T2(const Elem& x, const Elem& y) { return Term{&x, &y, 0}; }
inline Term TS(const Elem& x, uint64_t s) { return Term{&x, nullptr, s}; }

enum Kind : uint8_t { W, C, X };
struct Item {
    Kind kind;
    uint32_t cell;
    U256 v;
};
inline Item iw(const U256& v) { return Item{W, 0, v}; }
inline Item ic(const U256& v) { return Item{C, 0, v}; }
inline Item ix(uint32_t cell) { return Item{X, cell, U256{{0, 0, 0, 0}}}; }

struct Structure {   // what keygen needs; filled by the structure pass
    std::vector<std::vector<uint8_t>> q_enable;
    std::vector<uint8_t> q_lookup;
    std::vector<std::pair<uint32_t, uint32_t>> copies;
    std::vector<std::pair<uint32_t, uint32_t>> const_copies;   // (cell, constant index)
    std::unordered_map<U256, uint32_t, U256Hash> const_index;
    std::vector<U256> constants;
    std::vector<uint32_t> lookups;
    std::vector<uint64_t> rows;          // cells used per advice column (gate.., lookup..)
    // segment s >= 1 of run() (variable-base windows 1.., fixed-base windows 0.., final addition) starts with these
    // per-column row counters and these accumulator cells: what a worker thread needs to synthesise it independently
    struct Checkpoint {
        std::vector<uint64_t> rows;
        uint32_t acc_ids[8], facc_ids[8];   // x limbs 0-2, x native, y limbs 0-2, y native
        uint64_t cells_before, lookups_before;
    };
    std::vector<Checkpoint> checkpoints;
};

struct alignas(64) Builder {   // cache-line aligned: the synthesis workers hold adjacent copies (no false sharing between them)
    // parameters
    unsigned k, A, L, F, lb, LB;
    uint64_t n, u;
    bool selector_mode;
    unsigned top_bits, q_top_bits, carry_limbs, carry_bits;
    // state
    std::vector<uint64_t> rows;
    std::vector<uint64_t*> adv;       // A + L columns, 4 u64 per row, canonical
    Structure* st = nullptr;          // non-null: structure pass
    const FixedTables* tabs = nullptr;
    bool overflow = false;
    U256 P2[257];                     // powers of two
    U256 P2Z = U256{{0, 0, 0, 0}}, P2ONE = U256{{1, 0, 0, 0}};
    // region cursor
    unsigned cur_col = 0;
    uint64_t cur_row0 = 0, cur_len = 0;

    inline uint64_t* at(uint32_t cell) const { return adv[cell >> k] + 4 * (uint64_t)(cell & (n - 1)); }
    inline U256 val(uint32_t cell) const {
        U256 r;
        memcpy(r.l, at(cell), 32);
        return r;
    }

    void begin() {
        unsigned c = 0;
        for (unsigned i = 1; i < A; i++)
            if (rows[i] < rows[c]) c = i;
        cur_col = c;
        cur_row0 = rows[c];
        cur_len = 0;
    }
    inline uint32_t next_cell(uint64_t*& dst) {
        uint64_t row = cur_row0 + cur_len;
        if (row >= u) {
            overflow = true;
            row = u - 1;
        }
        cur_len++;
        dst = adv[cur_col] + 4 * row;
        return (uint32_t)(((uint64_t)cur_col << k) | row);
    }
    inline uint32_t put_w(const U256& v) {            // fresh witness (already reduced modulo r)
        uint64_t* dst;
        uint32_t cell = next_cell(dst);
        memcpy(dst, v.l, 32);
        return cell;
    }
    inline uint32_t put_c(const U256& v) {            // cell tied to the constants column
        uint64_t* dst;
        uint32_t cell = next_cell(dst);
        memcpy(dst, v.l, 32);
        if (st) {
            auto ins = st->const_index.emplace(v, (uint32_t)st->constants.size());
            if (ins.second) st->constants.push_back(v);
            st->const_copies.emplace_back(cell, ins.first->second);
        }
        return cell;
    }
    inline uint32_t put_xv(uint32_t src, const U256& v) {   // copy of an existing cell whose value the caller holds
        uint64_t* dst;
        uint32_t cell = next_cell(dst);
        memcpy(dst, v.l, 32);
        if (st) st->copies.emplace_back(src, cell);
        return cell;
    }
    inline uint32_t put_x(uint32_t src) { return put_xv(src, val(src)); }
    uint32_t put(const Item& it) { return it.kind == X ? put_x(it.cell) : (it.kind == C ? put_c(it.v) : put_w(it.v)); }
    inline void gate(uint64_t offset) {
        if (st && cur_row0 + offset < n) st->q_enable[cur_col][cur_row0 + offset] = 1;
    }
    void end() { rows[cur_col] = cur_row0 + cur_len; }

    uint64_t lk_cursor = 0;           // lookup cells so far (lookup-column mode): cell i -> column A + i mod L, row i div L
    void lookup(uint32_t cell) {
        if (selector_mode) {
            if (st) st->q_lookup[cell & (n - 1)] = 1;
            return;
        }
        // the witness pass copies the value into its lookup column now, while the cell is still in cache (a gather over the
        // finished columns afterwards costs a third of a millisecond, serial)
        const uint64_t i = lk_cursor++;
        if (st) st->lookups.push_back(cell);
        else if (L && i / L < u) memcpy(adv[A + i % L] + 4 * (i / L), at(cell), 32);
    }
    void equal(uint32_t a, uint32_t b) {
        if (st) st->copies.emplace_back(a, b);
    }

    // acc + x * y mod r for canonical values
    U256 muladd(const U256& acc, const U256& x, const U256& y) const {
        if (!(x.l[2] | x.l[3] | y.l[2] | y.l[3])) {      // limbs, bits, carries, small constants: the common case
            U256 r = mul128(low128(x), low128(y));
            if (!add_to(r, acc) && cmp(r, FR_MOD) < 0) return r;
        }
        uint64_t w[8];
        mul_wide(x, y, w);
        if (!(w[4] | w[5] | w[6] | w[7])) {
            U256 r = u256(w[0], w[1], w[2], w[3]);
            if (!add_to(r, acc) && cmp(r, FR_MOD) < 0) return r;
        }
        const Mont& f = moduli().fr;
        return f.add(f.mulmod(f.reduce(x), f.reduce(y)), acc);
    }
    inline U256 item_value(const Item& it) const { return it.kind == X ? val(it.cell) : it.v; }

    // cells [init, x0, y0, acc0, x1, y1, acc1, ...]; returns the cell of the final accumulator
    uint32_t chain(const Item& init, const Item* xs, const Item* ys, unsigned nterms, uint32_t* first_cell = nullptr) {
        begin();
        uint32_t last = put(init);
        if (first_cell) *first_cell = last;
        U256 acc = item_value(init);
        if (init.kind != X) acc = moduli().fr.reduce(acc);
        for (unsigned i = 0; i < nterms; i++) {
            acc = muladd(acc, item_value(xs[i]), item_value(ys[i]));
            put(xs[i]);
            put(ys[i]);
            last = put(iw(acc));
            gate(3 * i);
        }
        end();
        return last;
    }

    // the same chain, term by term (no Item arrays on the hot paths)
    U256 c_acc;
    unsigned c_terms = 0;
    uint32_t c_last = 0;
    inline void cbegin_c(const U256& init) {          // init already reduced modulo r
        begin();
        c_last = put_c(init);
        c_acc = init;
        c_terms = 0;
    }
    inline void cbegin_x(uint32_t cell) {
        begin();
        c_acc = val(cell);
        c_last = put_xv(cell, c_acc);
        c_terms = 0;
    }
    inline void cterm_xx(uint32_t a, uint32_t b) {
        const U256 av = val(a), bv = val(b);
        c_acc = muladd(c_acc, av, bv);
        put_xv(a, av);
        put_xv(b, bv);
        c_last = put_w(c_acc);
        gate(3 * c_terms++);
    }
    inline void cterm_vv(uint32_t a, const U256& av, uint32_t b, const U256& bv) {   // operands whose values the caller holds
        c_acc = muladd(c_acc, av, bv);
        put_xv(a, av);
        put_xv(b, bv);
        c_last = put_w(c_acc);
        gate(3 * c_terms++);
    }
    inline void cterm_vc(uint32_t a, const U256& av, const U256& cst) {
        c_acc = muladd(c_acc, av, cst);
        put_xv(a, av);
        put_c(cst);
        c_last = put_w(c_acc);
        gate(3 * c_terms++);
    }
    inline void cterm_xc(uint32_t a, const U256& cst) {
        const U256 av = val(a);
        c_acc = muladd(c_acc, av, cst);
        put_xv(a, av);
        put_c(cst);
        c_last = put_w(c_acc);
        gate(3 * c_terms++);
    }
    inline uint32_t cend() {
        end();
        return c_last;
    }

    // witness `value` < 2^bits (bits <= 128) as a fresh range-checked cell
    uint32_t range_limbs(u128 value, unsigned bits) {
        const unsigned kk = (bits + lb - 1) / lb, rem = bits % lb;
        const u128 m = ((u128)1 << lb) - 1;
        uint32_t limb_cells[16];
        uint32_t cell;
        begin();
        if (kk == 1) {
            cell = put_w(from128(value));
            limb_cells[0] = cell;
        } else {
            limb_cells[0] = put_w(from128(value & m));
            u128 acc = value & m;
            cell = limb_cells[0];
            for (unsigned i = 1; i < kk; i++) {
                u128 v = (value >> (lb * i)) & m;
                acc += v << (lb * i);
                limb_cells[i] = put_w(from128(v));
                put_c(P2[lb * i]);
                cell = put_w(from128(acc));
                gate(3 * (i - 1));
            }
        }
        end();
        for (unsigned i = 0; i < kk; i++) lookup(limb_cells[i]);
        if (rem) {
            u128 top = (value >> (lb * (kk - 1))) & m;
            begin();
            put_c(P2Z);
            put_xv(limb_cells[kk - 1], from128(top));
            put_c(P2[lb - rem]);
            uint32_t sh = put_w(from128(top << (lb - rem)));
            gate(0);
            end();
            lookup(sh);
        }
        return cell;
    }
    void assert_bit(uint32_t cell) {
        const U256 v = val(cell);
        begin();
        put_c(P2Z);
        put_xv(cell, v);
        put_xv(cell, v);
        put_xv(cell, v);
        gate(0);
        end();
    }

    void split(const U256& v, u128 out[3]) const {
        const u128 m = ((u128)1 << LB) - 1;
        out[0] = low128(v) & m;
        out[1] = low128(shr(v, LB)) & m;
        out[2] = low128(shr(v, 2 * LB));
    }
    uint32_t native_of(const uint32_t limbs[3]) {
        cbegin_x(limbs[0]);
        cterm_xc(limbs[1], P2[LB]);
        cterm_xc(limbs[2], P2[2 * LB]);
        return cend();
    }
    Elem new_elem(const U256& value, int top = -1) {
        Elem e;
        e.value = value;
        split(value, e.lv);
        const unsigned bits[3] = {LB, LB, top < 0 ? top_bits : (unsigned)top};
        for (int i = 0; i < 3; i++) {
            u128 v = e.lv[i];
            if (bits[i] < 128) v &= ((u128)1 << bits[i]) - 1;     // only differs for unsatisfiable inputs
            e.limbs[i] = range_limbs(v, bits[i]);
        }
        e.native = native_of(e.limbs);
        e.nat = c_acc;
        return e;
    }
    Elem const_elem(const U256& value) {
        Elem e;
        e.value = value;
        split(value, e.lv);
        begin();
        for (int i = 0; i < 3; i++) e.limbs[i] = put(ic(from128(e.lv[i])));
        e.nat = moduli().fr.reduce(value);
        e.native = put(ic(e.nat));
        end();
        return e;
    }

    // ---- the relation: sum pos - sum neg + k0 = 0 (mod m) -----------------------------------------------------
    struct ModConst {
        const Mont* f;
        U256 m;
        u128 ml[3];
        uint64_t minv[5];       // m^-1 mod 2^320
        uint64_t q0m[9];        // 2^258 * m
        U256 m_native;          // m mod r
        U256 q0m_native;        // 2^258 * m mod r
    };
    ModConst mc_p, mc_n;
    void init_mod(ModConst& c, const Mont& f) {
        c.f = &f;
        c.m = f.m;
        split(c.m, c.ml);
        // Newton iteration for m^-1 modulo 2^320 (5 limbs)
        uint64_t inv[5] = {0, 0, 0, 0, 0};
        uint64_t x = 1;
        for (int i = 0; i < 6; i++) x *= 2 - c.m.l[0] * x;
        inv[0] = x;
        uint64_t ml5[5] = {c.m.l[0], c.m.l[1], c.m.l[2], c.m.l[3], 0};
        for (int it = 0; it < 3; it++) {        // precision 64 -> 128 -> 256 -> 512 bits
            uint64_t t[5], two_minus[5];
            mul_low5(ml5, inv, t);
            // two_minus = 2 - t
            uint64_t borrow = 0;
            for (int i = 0; i < 5; i++) {
                u128 d = (u128)(i == 0 ? 2 : 0) - t[i] - borrow;
                two_minus[i] = (uint64_t)d;
                borrow = (uint64_t)(d >> 64) & 1;
            }
            uint64_t nx[5];
            mul_low5(inv, two_minus, nx);
            memcpy(inv, nx, 40);
        }
        memcpy(c.minv, inv, 40);
        memset(c.q0m, 0, sizeof c.q0m);
        // 2^258 = 2^(4*64 + 2)
        for (int i = 0; i < 4; i++) {
            c.q0m[4 + i] |= c.m.l[i] << 2;
            c.q0m[5 + i] |= c.m.l[i] >> 62;
        }
        const Mont& fr = moduli().fr;
        c.m_native = fr.reduce(c.m);
        U256 acc = u256(0);
        const U256 two64 = u256(0, 1);
        for (int i = 8; i >= 0; i--) acc = fr.add(fr.mulmod(acc, two64), fr.reduce(u256(c.q0m[i])));
        c.q0m_native = acc;
    }
    static void mul_low5(const uint64_t a[5], const uint64_t b[5], uint64_t out[5]) {
        uint64_t r[5] = {0, 0, 0, 0, 0};
        for (int i = 0; i < 5; i++) {
            u128 c = 0;
            for (int j = 0; i + j < 5; j++) { c += (u128)a[i] * b[j] + r[i + j]; r[i + j] = (uint64_t)c; c >>= 64; }
        }
        memcpy(out, r, 40);
    }
    static void acc_mul(uint64_t acc[9], const U256& a, const U256& b) {
        uint64_t w[8];
        mul_wide(a, b, w);
        u128 c = 0;
        for (int i = 0; i < 8; i++) { c += (u128)acc[i] + w[i]; acc[i] = (uint64_t)c; c >>= 64; }
        acc[8] += (uint64_t)c;
    }
    U256 term_limb_sum(const Term* ts, unsigned nt, unsigned i) const {
        U256 s = u256(0);
        for (unsigned t = 0; t < nt; t++) {
            if (ts[t].Y) {
                for (unsigned j = 0; j <= i; j++) add_to(s, mul128(ts[t].X->lv[j], ts[t].Y->lv[i - j]));
            } else {
                add_to(s, mul128(ts[t].X->lv[i], ts[t].scalar));
            }
        }
        return s;
    }

    void constrain(const ModConst& mc, const Term* pos, unsigned np, const Term* neg, unsigned nn, const U256& kpos_in, const U256& kneg) {
        // integer identity: POS + k+ + 2^258 m - NEG - k- = q' m
        uint64_t P9[9], N9[9];
        memcpy(P9, mc.q0m, sizeof P9);
        memset(N9, 0, sizeof N9);
        acc_mul(P9, kpos_in, u256(1));
        acc_mul(N9, kneg, u256(1));
        for (unsigned t = 0; t < np; t++) acc_mul(P9, pos[t].X->value, pos[t].Y ? pos[t].Y->value : u256(pos[t].scalar));
        for (unsigned t = 0; t < nn; t++) acc_mul(N9, neg[t].X->value, neg[t].Y ? neg[t].Y->value : u256(neg[t].scalar));
        uint64_t V5[5];
        {
            uint64_t borrow = 0;
            for (int i = 0; i < 5; i++) {
                u128 d = (u128)P9[i] - N9[i] - borrow;
                V5[i] = (uint64_t)d;
                borrow = (uint64_t)(d >> 64) & 1;
            }
        }
        uint64_t q5[5];
        mul_low5(V5, mc.minv, q5);          // exact quotient when m divides the value
        // q' < 2^259 does not fit 256 bits: cut its limbs straight from the 320-bit quotient
        Elem qe;
        {
            const u128 m = ((u128)1 << LB) - 1;
            U256 lo = u256(q5[0], q5[1], q5[2], q5[3]);
            qe.lv[0] = low128(lo) & m;
            qe.lv[1] = low128(shr(lo, LB)) & m;
            U256 hi = shr(lo, 2 * LB);
            add_to(hi, shl(u256(q5[4]), 256 - 2 * LB));
            qe.lv[2] = low128(hi) & (((u128)1 << q_top_bits) - 1);
            qe.value = u256(0);           // never used as an operand
            const unsigned bits[3] = {LB, LB, q_top_bits};
            for (int i = 0; i < 3; i++) qe.limbs[i] = range_limbs(qe.lv[i], bits[i]);
            qe.native = native_of(qe.limbs);
            qe.nat = c_acc;
        }
        // constant limbs: (k+ + 2^258 m) mod 2^(3 LB), k-
        uint64_t K9[9];
        memcpy(K9, mc.q0m, sizeof K9);
        acc_mul(K9, kpos_in, u256(1));
        U256 klow = u256(K9[0], K9[1], K9[2], K9[3]);   // only the low 3*LB <= 273 bits matter; limb 2 may need K9[4]
        u128 kp[3], kn[3];
        {
            const u128 m = ((u128)1 << LB) - 1;
            kp[0] = low128(klow) & m;
            kp[1] = low128(shr(klow, LB)) & m;
            // bits [2 LB, 3 LB) of the 576-bit constant
            U256 hi = shr(klow, 2 * LB);
            U256 k4 = shl(u256(K9[4], K9[5]), 256 - 2 * LB);
            add_to(hi, k4);
            kp[2] = low128(hi) & m;
        }
        split(kneg, kn);
        const U256 OFF = pow2(carry_bits - 1);
        const U256 OFFL = pow2(carry_bits - 1 + LB);
        uint32_t carry_prev = 0;
        const Mont& fr = moduli().fr;
        for (unsigned i = 0; i < 3; i++) {
            U256 linit = from128(kp[i]);
            add_to(linit, OFFL);
            cbegin_c(linit);
            limb_terms(pos, np, i);
            if (i > 0) cterm_xc(carry_prev, P2ONE);
            uint32_t lend = cend();
            U256 rinit = from128(kn[i]);
            if (i > 0) add_to(rinit, OFF);
            U256 rpart = rinit;
            add_to(rpart, term_limb_sum(neg, nn, i));
            for (unsigned j = 0; j <= i; j++) add_to(rpart, mul128(qe.lv[j], mc.ml[i - j]));
            U256 diff = c_acc;                       // value of the left chain's last cell
            sub_from(diff, rpart);
            u128 cw = low128(mask_bits(shr(diff, LB), carry_bits));
            uint32_t carry_cell = range_limbs(cw, carry_bits);
            cbegin_c(rinit);
            limb_terms(neg, nn, i);
            for (unsigned j = 0; j <= i; j++)
                if (mc.ml[i - j]) cterm_xc(qe.limbs[j], from128(mc.ml[i - j]));
            cterm_xc(carry_cell, P2[LB]);
            uint32_t rend = cend();
            equal(lend, rend);
            carry_prev = carry_cell;
        }
        // native: (k+ + 2^258 m) + sum pos = k- + sum neg + q' m   (mod r)
        const U256 knat = fr.add(mc.q0m_native, fr.reduce(kpos_in));   // (k+ + 2^258 m) mod r
        cbegin_c(knat);
        nat_terms(pos, np);
        uint32_t lend = cend();
        cbegin_c(fr.reduce(kneg));
        nat_terms(neg, nn);
        cterm_xc(qe.native, mc.m_native);
        uint32_t rend = cend();
        equal(lend, rend);
    }
    // operand values come from the Elems, never from the cells: an accumulator handed over from another worker thread's
    // segment may not have been written yet
    void limb_terms(const Term* ts, unsigned nt, unsigned i) {
        for (unsigned t = 0; t < nt; t++) {
            if (ts[t].Y) {
                for (unsigned j = 0; j <= i; j++)
                    cterm_vv(ts[t].X->limbs[j], from128(ts[t].X->lv[j]), ts[t].Y->limbs[i - j], from128(ts[t].Y->lv[i - j]));
            } else {
                cterm_vc(ts[t].X->limbs[i], from128(ts[t].X->lv[i]), u256(ts[t].scalar));
            }
        }
    }
    void nat_terms(const Term* ts, unsigned nt) {
        for (unsigned t = 0; t < nt; t++) {
            if (ts[t].Y) cterm_vv(ts[t].X->native, ts[t].X->nat, ts[t].Y->native, ts[t].Y->nat);
            else cterm_vc(ts[t].X->native, ts[t].X->nat, u256(ts[t].scalar));
        }
    }
    void constrain(const ModConst& mc, std::initializer_list<Term> pos, std::initializer_list<Term> neg, const U256& kpos = u256(0),
                   const U256& kneg = u256(0)) {
        constrain(mc, pos.begin(), (unsigned)pos.size(), neg.begin(), (unsigned)neg.size(), kpos, kneg);
    }

    // ---- field-level helpers ---------------------------------------------------------------------------------
    void assert_less_than(const Elem& a, const U256& bound) {
        U256 bm1 = bound;
        sub_from(bm1, u256(1));
        U256 d = bm1;
        sub_from(d, a.value);                 // wraps mod 2^256 when a >= bound (unsatisfiable)
        Elem de = new_elem(d);
        u128 b[3];
        split(bm1, b);
        uint32_t cin_cell = 0;
        u128 cin = 0;
        for (unsigned i = 0; i < 3; i++) {
            Item xs[2], ys[2];
            unsigned nt = 0;
            xs[nt] = ix(de.limbs[i]);
            ys[nt++] = ic(u256(1));
            if (i > 0) {
                xs[nt] = ix(cin_cell);
                ys[nt++] = ic(u256(1));
            }
            uint32_t lend = chain(ix(a.limbs[i]), xs, ys, nt);
            uint32_t rend;
            if (i < 2) {
                // cout = (a_i + d_i + cin - b_i) >> LB, a bit when the relation holds
                u128 s = a.lv[i] + de.lv[i] + cin;
                U256 coutv;
                if (s >= b[i]) {
                    coutv = from128((s - b[i]) >> LB);
                } else {                       // negative: unsatisfiable input, store the field representative
                    u128 neg = ((b[i] - s) + (((u128)1 << LB) - 1)) >> LB;
                    coutv = FR_MOD;
                    sub_from(coutv, from128(neg));
                }
                begin();
                uint32_t cc = put(iw(coutv));
                end();
                assert_bit(cc);
                Item x1[1] = {ix(cc)}, y1[1] = {ic(pow2(LB))};
                rend = chain(ic(from128(b[i])), x1, y1, 1);
                cin_cell = cc;
                cin = (s >= b[i]) ? ((s - b[i]) >> LB) : 0;
            } else {
                rend = chain(ic(from128(b[i])), nullptr, nullptr, 0);
            }
            equal(lend, rend);
        }
    }
    void assert_nonzero(const Elem& a) {
        Item xs[2] = {ix(a.limbs[1]), ix(a.limbs[2])};
        Item ys[2] = {ic(u256(1)), ic(u256(1))};
        uint32_t s = chain(ix(a.limbs[0]), xs, ys, 2);
        const Mont& fr = moduli().fr;
        U256 sv = val(s);
        U256 inv = is_zero(sv) ? u256(0) : fr.from(fr.inverse(fr.to(sv)));
        begin();
        put(ic(u256(0)));
        put(ix(s));
        put(iw(inv));
        put(ic(u256(1)));
        gate(0);
        end();
    }
    // canonical helpers over a modulus
    static U256 modmul(const Mont& f, const U256& a, const U256& b) { return f.mulmod(f.reduce(a), f.reduce(b)); }
    static U256 modinv(const Mont& f, const U256& a) {
        U256 r = f.reduce(a);
        return is_zero(r) ? u256(0) : f.from(f.inverse(f.to(r)));
    }
    Elem divide(const Elem& a, const Elem& b, const ModConst& mc) {
        const Mont& f = *mc.f;
        Elem uq = new_elem(modmul(f, a.value, modinv(f, b.value)));
        constrain(mc, {T2(uq, b)}, {TS(a, 1)});
        return uq;
    }

    // ---- bits, indicators, selection -------------------------------------------------------------------------
    void to_bits(uint32_t cell, unsigned nbits, std::vector<uint32_t>& out) {
        u128 v = low128(val(cell));
        begin();
        std::vector<uint32_t> bc(nbits);
        bc[0] = put(iw(u256((uint64_t)(v & 1))));
        u128 acc = v & 1;
        uint32_t last = bc[0];
        for (unsigned i = 1; i < nbits; i++) {
            u128 b = (v >> i) & 1;
            acc += b << i;
            bc[i] = put(iw(u256((uint64_t)b)));
            put(ic(pow2(i)));
            last = put(iw(from128(acc)));
            gate(3 * (i - 1));
        }
        end();
        equal(last, cell);
        for (unsigned i = 0; i < nbits; i++) assert_bit(bc[i]);
        out.insert(out.end(), bc.begin(), bc.end());
    }
    void indicator(const uint32_t bits_hi_to_lo[WINDOW], uint32_t ind[16]) {
        uint32_t cur[16], nxt[16];
        unsigned cnt = 2;
        uint32_t b = bits_hi_to_lo[0];
        uint64_t bv = val(b).l[0];
        begin();
        cur[0] = put(iw(u256(1 - bv)));
        put(ix(b));
        put(ic(u256(1)));
        put(ic(u256(1)));
        gate(0);
        end();
        cur[1] = b;
        for (unsigned lvl = 1; lvl < WINDOW; lvl++) {
            b = bits_hi_to_lo[lvl];
            bv = val(b).l[0];
            for (unsigned e = 0; e < cnt; e++) {
                uint64_t ev = val(cur[e]).l[0];
                begin();
                put(ic(u256(0)));
                put(ix(cur[e]));
                put(ix(b));
                uint32_t m = put(iw(u256(ev * bv)));
                gate(0);
                end();
                begin();
                uint32_t s = put(iw(u256(ev - ev * bv)));
                put(ix(m));
                put(ic(u256(1)));
                put(ix(cur[e]));
                gate(0);
                end();
                nxt[2 * e] = s;
                nxt[2 * e + 1] = m;
            }
            cnt *= 2;
            memcpy(cur, nxt, sizeof(uint32_t) * cnt);
        }
        memcpy(ind, cur, sizeof(uint32_t) * 16);
    }
    // sum_j ind_j * table_j; table entries are Elems (cells) or constants (tx[j])
    Elem select_elem(const uint32_t ind[16], const Elem* const* table, const U256* consts) {
        unsigned sel = 0;
        for (unsigned j = 0; j < 16; j++)
            if (val(ind[j]).l[0]) sel = j;
        Elem e;
        for (unsigned i = 0; i < 3; i++) {
            cbegin_c(P2Z);
            for (unsigned j = 0; j < 16; j++) {
                if (consts) {
                    u128 lv[3];
                    split(consts[j], lv);
                    cterm_xc(ind[j], from128(lv[i]));
                } else {
                    cterm_xx(ind[j], table[j]->limbs[i]);
                }
            }
            e.limbs[i] = cend();
        }
        e.native = native_of(e.limbs);
        e.nat = c_acc;
        e.value = consts ? consts[sel] : table[sel]->value;
        split(e.value, e.lv);
        return e;
    }

    // ---- curve ops over CRT elements ----------------------------------------------------------------------------
    struct EPt { Elem x, y; };
    // inverses of the denominators of every addition / doubling, in the order run() performs them, from one batch
    // inversion over a projective pre-pass (precompute_denominators); empty = invert one by one
    std::vector<U256> dinv_queue;
    size_t dinv_next = 0;
    // affine coordinates (canonical) of the variable-base accumulator after window wi and of the fixed-base accumulator
    // after window wi, from the same pre-pass: lets a worker thread start in the middle of run()
    std::vector<U256> traj_acc_x, traj_acc_y, traj_facc_x, traj_facc_y;
    U256 denominator_inverse(const U256& d) {
        if (dinv_next < dinv_queue.size()) {
            const U256& c = dinv_queue[dinv_next++];
            if (moduli().fp.mulmod(c, d) == u256(1)) return c;
            dinv_queue.clear();                     // trajectory diverged (degenerate input): fall back
        }
        return modinv(moduli().fp, d);
    }
    U256 psub(const U256& a, const U256& b) const { return moduli().fp.sub(a, b); }   // canonical values < p
    U256 pmul(const U256& a, const U256& b) const { return moduli().fp.mulmod(a, b); }
    U256 pred(const U256& a) const { return moduli().fp.reduce(a); }

    EPt ec_add(const EPt& P, const EPt& Q, bool strict = false) {
        const U256 x1 = pred(P.x.value), y1 = pred(P.y.value), x2 = pred(Q.x.value), y2 = pred(Q.y.value);
        U256 dx = psub(x2, x1);
        U256 dxinv = denominator_inverse(dx);
        if (strict) {
            Elem t = new_elem(dxinv);
            constrain(mc_p, {T2(t, Q.x)}, {T2(t, P.x)}, u256(0), u256(1));
        }
        U256 lamv = pmul(psub(y2, y1), dxinv);
        Elem lam = new_elem(lamv);
        constrain(mc_p, {T2(lam, Q.x), TS(P.y, 1)}, {T2(lam, P.x), TS(Q.y, 1)});
        U256 x3v = psub(psub(pmul(lamv, lamv), x1), x2);
        EPt out;
        out.x = new_elem(x3v);
        constrain(mc_p, {T2(lam, lam)}, {TS(P.x, 1), TS(Q.x, 1), TS(out.x, 1)});
        U256 y3v = psub(pmul(lamv, psub(x1, x3v)), y1);
        out.y = new_elem(y3v);
        constrain(mc_p, {T2(lam, P.x)}, {T2(lam, out.x), TS(P.y, 1), TS(out.y, 1)});
        return out;
    }
    EPt ec_double(const EPt& P) {
        const Mont& f = moduli().fp;
        const U256 x = pred(P.x.value), y = pred(P.y.value);
        U256 num = psub(pmul(u256(3), pmul(x, x)), u256(3));
        U256 lamv = pmul(num, denominator_inverse(f.add(y, y)));
        Elem lam = new_elem(lamv);
        constrain(mc_p, {T2(lam, P.y), T2(lam, P.y)}, {T2(P.x, P.x), T2(P.x, P.x), T2(P.x, P.x)}, u256(3));
        U256 x3v = psub(psub(pmul(lamv, lamv), x), x);
        EPt out;
        out.x = new_elem(x3v);
        constrain(mc_p, {T2(lam, lam)}, {TS(P.x, 2), TS(out.x, 1)});
        U256 y3v = psub(pmul(lamv, psub(x, x3v)), y);
        out.y = new_elem(y3v);
        constrain(mc_p, {T2(lam, P.x)}, {T2(lam, out.x), TS(P.y, 1), TS(out.y, 1)});
        return out;
    }
    void assert_on_curve(const EPt& P) {
        Elem x2 = new_elem(pmul(pred(P.x.value), pred(P.x.value)));
        constrain(mc_p, {T2(P.x, P.x)}, {TS(x2, 1)});
        constrain(mc_p, {T2(P.y, P.y), TS(P.x, 3)}, {T2(x2, P.x)}, u256(0), P256_B);
    }
    // 4-bit window indicators of the 3*LB bits of u, most significant window first
    void scalar_windows(const Elem& uq, std::vector<uint32_t>& inds) {
        std::vector<uint32_t> bits;
        bits.reserve(3 * LB + WINDOW);
        for (int i = 0; i < 3; i++) to_bits(uq.limbs[i], LB, bits);
        while (bits.size() % WINDOW) {
            begin();
            bits.push_back(put(ic(u256(0))));
            end();
        }
        const unsigned nw = (unsigned)bits.size() / WINDOW;
        inds.resize(16 * nw);
        for (unsigned wi = 0; wi < nw; wi++) {
            unsigned w = nw - 1 - wi;
            uint32_t hl[WINDOW];
            for (unsigned j = 0; j < WINDOW; j++) hl[j] = bits[WINDOW * w + (WINDOW - 1 - j)];
            indicator(hl, &inds[16 * wi]);
        }
    }

    // ---- projective pre-pass: all denominators of run()'s curve operations with ONE field inversion ---------------------
    struct Jac { U256 X, Y, Z; };   // Montgomery form
    static Jac jac_dbl(const Mont& f, const Jac& p) {          // a = -3 (dbl-2001-b)
        U256 delta = f.mul(p.Z, p.Z), gamma = f.mul(p.Y, p.Y), beta = f.mul(p.X, gamma);
        U256 t = f.mul(f.sub(p.X, delta), f.add(p.X, delta));
        U256 alpha = f.add(f.add(t, t), t);
        U256 b2 = f.add(beta, beta), b4 = f.add(b2, b2), b8 = f.add(b4, b4);
        Jac r;
        r.X = f.sub(f.mul(alpha, alpha), b8);
        U256 yz = f.add(p.Y, p.Z);
        r.Z = f.sub(f.sub(f.mul(yz, yz), gamma), delta);
        U256 g2 = f.mul(gamma, gamma);
        U256 g4 = f.add(g2, g2);
        g4 = f.add(g4, g4);
        U256 g8 = f.add(g4, g4);
        r.Y = f.sub(f.mul(alpha, f.sub(b4, r.X)), g8);
        return r;
    }
    static bool jac_add(const Mont& f, const Jac& p, const Jac& q, Jac& r) {   // add-2007-bl; false when x1 == x2
        U256 z1z1 = f.mul(p.Z, p.Z), z2z2 = f.mul(q.Z, q.Z);
        U256 u1 = f.mul(p.X, z2z2), u2 = f.mul(q.X, z1z1);
        U256 s1 = f.mul(f.mul(p.Y, q.Z), z2z2), s2 = f.mul(f.mul(q.Y, p.Z), z1z1);
        U256 h = f.sub(u2, u1);
        if (is_zero(h)) return false;
        U256 h2 = f.add(h, h);
        U256 i = f.mul(h2, h2), j = f.mul(h, i);
        U256 rr = f.sub(s2, s1);
        rr = f.add(rr, rr);
        U256 v = f.mul(u1, i);
        r.X = f.sub(f.sub(f.mul(rr, rr), j), f.add(v, v));
        U256 s1j = f.mul(s1, j);
        r.Y = f.sub(f.mul(rr, f.sub(v, r.X)), f.add(s1j, s1j));
        U256 zz = f.add(p.Z, q.Z);
        r.Z = f.mul(f.sub(f.sub(f.mul(zz, zz), z1z1), z2z2), h);
        return !is_zero(r.Z);
    }
    static void batch_invert(const Mont& f, std::vector<U256>& v) {   // Montgomery in / out; entries must be non-zero
        const size_t n = v.size();
        if (!n) return;
        std::vector<U256> pre(n);
        U256 acc = f.one;
        for (size_t i = 0; i < n; i++) {
            pre[i] = acc;
            acc = f.mul(acc, v[i]);
        }
        U256 inv = f.inverse(acc);
        for (size_t i = n; i-- > 0;) {
            U256 t = f.mul(inv, pre[i]);
            inv = f.mul(inv, v[i]);
            v[i] = t;
        }
    }
    // digits of the 3*LB-bit scalar, most significant 4-bit window first
    void window_digits(const U256& uv, std::vector<unsigned>& d) const {
        const unsigned nw = (3 * LB + WINDOW - 1) / WINDOW;
        d.resize(nw);
        for (unsigned wi = 0; wi < nw; wi++) {
            unsigned w = nw - 1 - wi;
            d[wi] = 4 * w < 256 ? (unsigned)(shr(uv, 4 * w).l[0] & 15) : 0;
        }
    }
    bool precompute_denominators(const U256& pkx, const U256& pky, const U256& r, const U256& s, const U256& msghash) {
        const Mont& f = moduli().fp;
        const Mont& fn = moduli().fn;
        dinv_queue.clear();
        dinv_next = 0;
        if (cmp(pkx, P256_P) >= 0 || cmp(pky, P256_P) >= 0) return false;
        U256 sinv = modinv(fn, s);
        U256 u1 = modmul(fn, msghash, sinv), u2 = modmul(fn, r, sinv);
        std::vector<unsigned> d1, d2;
        window_digits(u1, d1);
        window_digits(u2, d2);
        const unsigned nw = (unsigned)d1.size();
        struct Op { int a, b, out; };            // b < 0: doubling of a
        std::vector<Jac> pts;
        std::vector<Op> ops;
        pts.reserve(6 * nw + 40);
        ops.reserve(6 * nw + 40);
        auto affine = [&](const U256& x, const U256& y) {
            pts.push_back(Jac{f.to(x), f.to(y), f.one});
            return (int)pts.size() - 1;
        };
        auto add = [&](int a, int b) {
            Jac rj;
            if (!jac_add(f, pts[a], pts[b], rj)) return -1;
            pts.push_back(rj);
            ops.push_back(Op{a, b, (int)pts.size() - 1});
            return (int)pts.size() - 1;
        };
        auto dbl = [&](int a) {
            pts.push_back(jac_dbl(f, pts[a]));
            ops.push_back(Op{a, -1, (int)pts.size() - 1});
            return (int)pts.size() - 1;
        };
        const int pk = affine(pkx, pky);
        int table[16];
        table[0] = affine(OFF_VAR_X, OFF_VAR_Y);
        for (unsigned j = 1; j < 16; j++)
            if ((table[j] = add(table[j - 1], pk)) < 0) return false;
        std::vector<int> acc_at(nw), facc_at(nw);
        int acc = table[d2[0]];
        acc_at[0] = acc;
        for (unsigned wi = 1; wi < nw; wi++) {
            for (unsigned i = 0; i < WINDOW; i++) acc = dbl(acc);
            if ((acc = add(acc, table[d2[wi]])) < 0) return false;
            acc_at[wi] = acc;
        }
        int facc = affine(tabs->start_x, tabs->start_y);
        for (unsigned wi = 0; wi < nw; wi++) {
            unsigned w = nw - 1 - wi;
            int t = affine(tabs->x[16 * w + d1[wi]], tabs->y[16 * w + d1[wi]]);
            if ((facc = add(facc, t)) < 0) return false;
            facc_at[wi] = facc;
        }
        if (add(acc, facc) < 0) return false;
        // affine coordinates of every point: x = X / Z^2, y = Y / Z^3
        std::vector<U256> zi(pts.size());
        for (size_t i = 0; i < pts.size(); i++) {
            if (is_zero(pts[i].Z)) return false;
            zi[i] = pts[i].Z;
        }
        batch_invert(f, zi);
        std::vector<U256> ax(pts.size()), ay(pts.size());
        for (size_t i = 0; i < pts.size(); i++) {
            U256 z2 = f.mul(zi[i], zi[i]);
            ax[i] = f.mul(pts[i].X, z2);
            ay[i] = f.mul(pts[i].Y, f.mul(z2, zi[i]));
        }
        std::vector<U256> den(ops.size());
        for (size_t i = 0; i < ops.size(); i++) {
            den[i] = ops[i].b < 0 ? f.add(ay[ops[i].a], ay[ops[i].a]) : f.sub(ax[ops[i].b], ax[ops[i].a]);
            if (is_zero(den[i])) return false;
        }
        batch_invert(f, den);
        dinv_queue.resize(ops.size());
        for (size_t i = 0; i < ops.size(); i++) dinv_queue[i] = f.from(den[i]);
        traj_acc_x.resize(nw); traj_acc_y.resize(nw); traj_facc_x.resize(nw); traj_facc_y.resize(nw);
        for (unsigned wi = 0; wi < nw; wi++) {
            traj_acc_x[wi] = f.from(ax[acc_at[wi]]); traj_acc_y[wi] = f.from(ay[acc_at[wi]]);
            traj_facc_x[wi] = f.from(ax[facc_at[wi]]); traj_facc_y[wi] = f.from(ay[facc_at[wi]]);
        }
        return true;
    }

    // ---- the circuit (mirrors oracle/ecdsa_circuit.py: synthesize) -------------------------------------------------
    // Everything the window loops read but do not write: produced by the prologue, shared (read-only) by the workers.
    struct RunCtx {
        std::vector<EPt> table;
        std::vector<uint32_t> inds_var, inds_fix;
        Elem r_e;
        U256 r, s;
        unsigned nw = 0;
    };
    unsigned num_segments(unsigned nw) const { return 2 * nw + 1; }   // 0 prologue | 1..nw-1 variable | nw..2nw-1 fixed | 2nw final

    // ecdsa_p256.rs:139-177 (load m, r, s, the public key) through the first variable-base window
    void prologue(const U256& pkx, const U256& pky, const U256& r, const U256& s, const U256& msghash, RunCtx& rc, EPt& acc, EPt& facc) {
        Elem m_e = new_elem(msghash), r_e = new_elem(r), s_e = new_elem(s);
        EPt pk{new_elem(pkx), new_elem(pky)};
        assert_on_curve(pk);
        assert_nonzero(r_e);
        assert_less_than(r_e, P256_N);
        assert_nonzero(s_e);
        assert_less_than(s_e, P256_N);
        Elem u1 = divide(m_e, s_e, mc_n);
        Elem u2 = divide(r_e, s_e, mc_n);
        assert_less_than(u1, P256_N);
        assert_less_than(u2, P256_N);
        rc.r_e = r_e;
        rc.r = r;
        rc.s = s;
        // variable base: table T[d] = d * PK + B2
        rc.table.resize(16);
        rc.table[0] = EPt{const_elem(OFF_VAR_X), const_elem(OFF_VAR_Y)};
        for (unsigned j = 1; j < 16; j++) rc.table[j] = ec_add(rc.table[j - 1], pk);
        // window indicators of both scalars and the fixed-base start point, then the first window: acc = T[top digit]
        scalar_windows(u2, rc.inds_var);
        scalar_windows(u1, rc.inds_fix);
        rc.nw = (unsigned)rc.inds_var.size() / 16;
        facc = EPt{const_elem(tabs->start_x), const_elem(tabs->start_y)};
        acc = select_point(rc, 0);
    }
    EPt select_point(const RunCtx& rc, unsigned wi) {
        const Elem* tx[16];
        const Elem* ty[16];
        for (unsigned j = 0; j < 16; j++) {
            tx[j] = &rc.table[j].x;
            ty[j] = &rc.table[j].y;
        }
        EPt sel;
        sel.x = select_elem(&rc.inds_var[16 * wi], tx, nullptr);
        sel.y = select_elem(&rc.inds_var[16 * wi], ty, nullptr);
        return sel;
    }
    static void ids_of(const EPt& p, uint32_t ids[8]) {
        for (int i = 0; i < 3; i++) { ids[i] = p.x.limbs[i]; ids[4 + i] = p.y.limbs[i]; }
        ids[3] = p.x.native;
        ids[7] = p.y.native;
    }
    Elem elem_at(const uint32_t ids[4], const U256& value) const {
        Elem e;
        for (int i = 0; i < 3; i++) e.limbs[i] = ids[i];
        e.native = ids[3];
        e.value = value;
        split(value, e.lv);
        e.nat = moduli().fr.reduce(value);
        return e;
    }
    // segments [s0, s1) of run(), in order; acc / facc are the accumulators at the start of s0 and at the end of s1 - 1
    void run_segments(const RunCtx& rc, unsigned s0, unsigned s1, EPt& acc, EPt& facc, bool* sig_ok) {
        const unsigned nw = rc.nw;
        for (unsigned sgm = s0; sgm < s1; sgm++) {
            if (st) {
                Structure::Checkpoint cp;
                cp.rows = rows;
                ids_of(acc, cp.acc_ids);
                ids_of(facc, cp.facc_ids);
                cp.cells_before = 0;
                for (uint64_t r : rows) cp.cells_before += r;
                cp.lookups_before = lk_cursor;
                if (st->checkpoints.size() <= sgm) st->checkpoints.resize(sgm + 1);
                st->checkpoints[sgm] = cp;
            }
            if (sgm < nw) {                       // variable base: acc = 16 * acc + T[window]
                EPt sel = select_point(rc, sgm);
                for (unsigned i = 0; i < WINDOW; i++) acc = ec_double(acc);
                acc = ec_add(acc, sel);
            } else if (sgm < 2 * nw) {            // fixed base: facc += Tab[w][window]
                const unsigned wi = sgm - nw, w = nw - 1 - wi;
                EPt sel;
                sel.x = select_elem(&rc.inds_fix[16 * wi], nullptr, &tabs->x[16 * w]);
                sel.y = select_elem(&rc.inds_fix[16 * wi], nullptr, &tabs->y[16 * w]);
                facc = ec_add(facc, sel);
            } else {                              // R = u1*G + u2*PK (strict), R.x == r limb by limb
                EPt R = ec_add(acc, facc, true);
                bool ok = true;
                for (int i = 0; i < 3; i++) {
                    equal(R.x.limbs[i], rc.r_e.limbs[i]);
                    ok = ok && R.x.lv[i] == rc.r_e.lv[i];
                }
                // r, s in [1, n-1] (the copy constraint above is necessary, not sufficient)
                ok = ok && !is_zero(rc.r) && !is_zero(rc.s) && cmp(rc.r, P256_N) < 0 && cmp(rc.s, P256_N) < 0;
                if (sig_ok) *sig_ok = ok;
            }
        }
    }
    // index into dinv_queue of the first curve operation of segment sgm (15 table additions come first)
    static size_t first_op_of(unsigned sgm, unsigned nw) {
        if (sgm < nw) return 15 + (size_t)(sgm - 1) * (WINDOW + 1);
        const size_t var_ops = 15 + (size_t)(nw - 1) * (WINDOW + 1);
        return var_ops + (sgm - nw);
    }
    // the whole circuit on the calling thread (mirrors oracle/ecdsa_circuit.py: synthesize)
    void run(const U256& pkx, const U256& pky, const U256& r, const U256& s, const U256& msghash, bool* sig_ok) {
        RunCtx rc;
        EPt acc, facc;
        prologue(pkx, pky, r, s, msghash, rc, acc, facc);
        run_segments(rc, 1, num_segments(rc.nw), acc, facc, sig_ok);
    }
};

}  // namespace

struct zkw_ecdsa_circuit {
    zkw_circuit_params params;
    zkw_circuit_shape shape;
    Structure st;
    FixedTables tabs;
    bool fits = true;
};

namespace {

void setup_builder(Builder& b, const zkw_ecdsa_circuit* c) {
    const zkw_circuit_params& p = c->params;
    b.k = p.degree;
    b.A = p.num_advice;
    b.selector_mode = p.num_advice == 1;
    b.L = b.selector_mode ? 0 : p.num_lookup_advice;
    b.F = p.num_fixed;
    b.lb = p.lookup_bits;
    b.LB = p.limb_bits;
    b.n = 1ull << p.degree;
    b.u = b.n - (c->shape.blinding_factors + 1);
    b.top_bits = 256 - 2 * b.LB;
    b.q_top_bits = Q_OFFSET_BITS + 1 - 2 * b.LB;
    b.carry_limbs = (b.LB + 6 + b.lb - 1) / b.lb;
    b.carry_bits = b.carry_limbs * b.lb;
    b.rows.assign(b.A, 0);
    b.tabs = &c->tabs;
    for (unsigned i = 0; i < 256; i++) b.P2[i] = pow2(i);
    b.P2[256] = u256(0);
    b.init_mod(b.mc_p, moduli().fp);
    b.init_mod(b.mc_n, moduli().fn);
}

// Write the cache lines of [p, p + bytes) back to memory (they stay cached, clean).  The assigned cells are shipped to the
// device by DMA right after synthesis; lines still Modified in the private caches of the eight worker cores make that copy
// crawl - measured on the GPU box (tools/h2d_dirty_test.py): 12.8 MB page-locked, at rest 0.25 ms, just written by eight
// threads 1.1-1.2 ms (1.9 ms in the prover), written and clwb'ed by the same threads 0.25 ms at no measurable cost to them.
#if defined(__x86_64__)
__attribute__((target("clwb"))) static void wb_clwb(const char* p, const char* e) {
    for (; p < e; p += 64) _mm_clwb((void*)p);
    _mm_sfence();
}
__attribute__((target("clflushopt"))) static void wb_clflushopt(const char* p, const char* e) {
    for (; p < e; p += 64) _mm_clflushopt((void*)p);
    _mm_sfence();
}
static int writeback_kind() {
    static const int kind = [] {
        if (getenv("ZKW_SYNTH_NO_WRITEBACK")) return 0;
        unsigned a = 0, b = 0, c = 0, d = 0;
        if (!__get_cpuid_count(7, 0, &a, &b, &c, &d)) return 0;
        return ((b >> 24) & 1u) ? 2 : (((b >> 23) & 1u) ? 1 : 0);   // CLWB, else CLFLUSHOPT, else leave the lines alone
    }();
    return kind;
}
static void writeback_lines(const void* ptr, size_t bytes) {
    const int kind = writeback_kind();
    if (!kind || !bytes) return;
    const char* p = (const char*)((uintptr_t)ptr & ~(uintptr_t)63);
    const char* e = (const char*)ptr + bytes;
    if (kind == 2) wb_clwb(p, e); else wb_clflushopt(p, e);
}
#else
static void writeback_lines(const void*, size_t) {}
#endif

// worker threads of one synthesis: ZKW_SYNTH_THREADS, default min(8, hardware threads / 2).  Measured on the GPU box's 16-core
// host (ZKW_SYNTH_TIMING=1, tools/synth_timing.py): 0.57 ms serial (denominator pre-pass 0.32, prologue 0.25) + 3.0 ms / threads
// + ~0.2 ms of thread start/join: 3.9 ms with one thread, 1.6-1.9 with four, 1.25-1.4 with eight.
unsigned synth_threads() {
    const char* e = getenv("ZKW_SYNTH_THREADS");
    if (e && atoi(e) > 0) return (unsigned)std::min(atoi(e), 16);
    static const unsigned hw = std::thread::hardware_concurrency();
    return std::max(1u, std::min(8u, hw / 2));
}

U256 load_le(const uint8_t b[32]) {
    U256 r;
    memcpy(r.l, b, 32);   // little-endian bytes on a little-endian host
    return r;
}

}  // namespace

extern "C" int zkw_ecdsa_circuit_new(const zkw_circuit_params* p, zkw_ecdsa_circuit** out) {
    if (!p || !out) return ZKW_ERR_INVALID;
    if (p->num_limbs != 3 || p->degree < 10 || p->degree > 24 || p->num_advice == 0 || p->num_fixed == 0 || p->lookup_bits < 8 ||
        p->lookup_bits > 24 || p->lookup_bits >= p->degree + 1 || p->limb_bits < 86 || p->limb_bits > 96 ||
        (p->num_advice > 1 && p->num_lookup_advice == 0) || (uint64_t)(p->num_advice + p->num_lookup_advice) << p->degree > 0xFFFFFFFFull)
        return ZKW_ERR_INVALID;
    zkw_ecdsa_circuit* c = new zkw_ecdsa_circuit();
    c->params = *p;
    const bool sel = p->num_advice == 1;
    zkw_circuit_shape& s = c->shape;
    s.k = p->degree;
    s.num_advice = p->num_advice;
    s.num_lookup_advice = sel ? 0 : p->num_lookup_advice;
    s.num_fixed = p->num_fixed;
    s.blinding_factors = 6;
    s.cs_degree = sel ? 5 : 4;
    s.ext_k = p->degree + 2;
    s.reserved = 0;
    c->tabs.build(p->limb_bits);
    // structure pass on a fixed valid assertion (sk = 1, nonce = 1, m = 1): the layout does not depend on the values
    Builder b;
    setup_builder(b, c);
    const uint64_t n = b.n;
    std::vector<std::vector<uint64_t>> cols(b.A + b.L, std::vector<uint64_t>(4 * n, 0));
    for (auto& col : cols) b.adv.push_back(col.data());
    c->st.q_enable.assign(b.A, std::vector<uint8_t>(n, 0));
    c->st.q_lookup.assign(n, 0);
    b.st = &c->st;
    const Mont& fn = moduli().fn;
    U256 r = fn.reduce(P256_GX);
    U256 s1 = fn.add(u256(1), r);    // s = k^-1 (m + r sk) with k = sk = m = 1
    b.run(P256_GX, P256_GY, r, s1, u256(1), nullptr);
    c->st.rows.assign(b.A + b.L, 0);
    for (unsigned i = 0; i < b.A; i++) c->st.rows[i] = b.rows[i];
    if (!b.selector_mode) {
        const uint64_t nl = c->st.lookups.size();
        if (nl > (uint64_t)b.L * b.u) b.overflow = true;
        for (unsigned l = 0; l < b.L; l++) c->st.rows[b.A + l] = std::min<uint64_t>(b.u, (nl + b.L - 1 - l) / b.L);
    }
    if (c->st.constants.size() > (uint64_t)b.F * b.u) b.overflow = true;
    c->fits = !b.overflow;
    *out = c;
    return c->fits ? ZKW_OK : ZKW_ERR_UNSUPPORTED;
}

extern "C" void zkw_ecdsa_circuit_free(zkw_ecdsa_circuit* c) { delete c; }

extern "C" int zkw_ecdsa_circuit_shape(const zkw_ecdsa_circuit* c, zkw_circuit_shape* out) {
    if (!c || !out) return ZKW_ERR_INVALID;
    *out = c->shape;
    return ZKW_OK;
}

extern "C" int zkw_ecdsa_circuit_rows(const zkw_ecdsa_circuit* c, size_t* rows_out, uint64_t* stats_out) {
    if (!c || !rows_out) return ZKW_ERR_INVALID;
    for (size_t i = 0; i < c->st.rows.size(); i++) rows_out[i] = (size_t)c->st.rows[i];
    if (stats_out) {
        uint64_t cells = 0;
        for (unsigned i = 0; i < c->params.num_advice; i++) cells += c->st.rows[i];
        uint64_t lk = 0;
        if (c->shape.num_lookup_advice == 0)
            for (uint8_t q : c->st.q_lookup) lk += q;
        else
            lk = c->st.lookups.size();
        stats_out[0] = cells;
        stats_out[1] = lk;
        stats_out[2] = c->st.constants.size();
        stats_out[3] = c->st.copies.size() + c->st.const_copies.size();
    }
    return ZKW_OK;
}

extern "C" int zkw_ecdsa_circuit_fixed(const zkw_ecdsa_circuit* c, uint64_t* const* fixed_out) {
    if (!c || !fixed_out || !c->fits) return ZKW_ERR_INVALID;
    const zkw_circuit_shape& s = c->shape;
    const uint64_t n = 1ull << s.k, u = n - (s.blinding_factors + 1);
    const unsigned F = s.num_fixed, A = s.num_advice;
    const unsigned ncols = F + 1 + A + (s.num_lookup_advice == 0 ? 1 : 0);
    for (unsigned i = 0; i < ncols; i++) {
        if (!fixed_out[i]) return ZKW_ERR_INVALID;
        memset(fixed_out[i], 0, 32 * n);
    }
    for (size_t idx = 0; idx < c->st.constants.size(); idx++) memcpy(fixed_out[idx % F] + 4 * (idx / F), c->st.constants[idx].l, 32);
    uint64_t T = 1ull << c->params.lookup_bits;
    if (T > u) T = u;
    for (uint64_t i = 0; i < T; i++) fixed_out[F][4 * i] = i;
    for (unsigned a = 0; a < A; a++)
        for (uint64_t i = 0; i < n; i++) fixed_out[F + 1 + a][4 * i] = c->st.q_enable[a][i];
    if (s.num_lookup_advice == 0)
        for (uint64_t i = 0; i < n; i++) fixed_out[F + 1 + A][4 * i] = c->st.q_lookup[i];
    return ZKW_OK;
}

extern "C" int zkw_ecdsa_circuit_permutation(const zkw_ecdsa_circuit* c, uint32_t* const* mapping_out) {
    if (!c || !mapping_out || !c->fits) return ZKW_ERR_INVALID;
    const zkw_circuit_shape& s = c->shape;
    const unsigned k = s.k, F = s.num_fixed, A = s.num_advice, L = s.num_lookup_advice;
    const uint64_t n = 1ull << k;
    const unsigned ncols = F + A + L;
    // union-find over permutation-column cell ids (col * n + row); advice cells are offset by the F constant columns
    std::vector<uint32_t> parent((size_t)ncols * n);
    for (size_t i = 0; i < parent.size(); i++) parent[i] = (uint32_t)i;
    auto find = [&](uint32_t x) {
        while (parent[x] != x) {
            parent[x] = parent[parent[x]];
            x = parent[x];
        }
        return x;
    };
    auto unite = [&](uint32_t a, uint32_t b) {
        uint32_t ra = find(a), rb = find(b);
        if (ra != rb) parent[std::max(ra, rb)] = std::min(ra, rb);
    };
    const uint32_t base = (uint32_t)(F * n);
    for (auto& pr : c->st.copies) unite(base + pr.first, base + pr.second);
    for (auto& pr : c->st.const_copies) unite(base + pr.first, (uint32_t)((pr.second % F) * n + pr.second / F));
    if (L) {   // lookup cells: cell i -> column A + i mod L, row i div L (the analogue of fp_chip.finalize, ecdsa_p256.rs:193-195)
        for (size_t i = 0; i < c->st.lookups.size(); i++)
            unite(base + c->st.lookups[i], base + (uint32_t)(((uint64_t)(A + i % L) << k) | (i / L)));
    }
    // cycles: members of each class in increasing id order, closed
    std::vector<uint32_t> last((size_t)ncols * n, 0xFFFFFFFFu), first((size_t)ncols * n, 0xFFFFFFFFu);
    for (unsigned col = 0; col < ncols; col++)
        for (uint64_t r = 0; r < n; r++) {
            mapping_out[col][2 * r] = col;
            mapping_out[col][2 * r + 1] = (uint32_t)r;
        }
    for (uint32_t x = 0; x < (uint32_t)parent.size(); x++) {
        uint32_t root = find(x);
        if (first[root] == 0xFFFFFFFFu) {
            first[root] = x;
            last[root] = x;
        } else {
            uint32_t p = last[root];
            mapping_out[p >> k][2 * (p & (n - 1))] = x >> k;
            mapping_out[p >> k][2 * (p & (n - 1)) + 1] = x & (uint32_t)(n - 1);
            last[root] = x;
        }
    }
    for (uint32_t root = 0; root < (uint32_t)parent.size(); root++) {
        if (first[root] != 0xFFFFFFFFu && last[root] != first[root]) {
            uint32_t p = last[root], x = first[root];
            mapping_out[p >> k][2 * (p & (n - 1))] = x >> k;
            mapping_out[p >> k][2 * (p & (n - 1)) + 1] = x & (uint32_t)(n - 1);
        }
    }
    return ZKW_OK;
}

extern "C" int zkw_ecdsa_synthesize(const zkw_ecdsa_circuit* c, const uint8_t pubkey_x[32], const uint8_t pubkey_y[32], const uint8_t r[32],
                                    const uint8_t s[32], const uint8_t msg_hash[32], uint64_t* const* advice_out, size_t* rows_out,
                                    int* signature_ok) {
    if (!c || !c->fits || !pubkey_x || !pubkey_y || !r || !s || !msg_hash || !advice_out) return ZKW_ERR_INVALID;
    Builder b;
    auto T0 = std::chrono::steady_clock::now();
    const bool timing = getenv("ZKW_SYNTH_TIMING") != nullptr;   // development aid: phase and per-thread times on stderr
    auto lap = [&](const char* what) { if (timing) { auto t = std::chrono::steady_clock::now(); fprintf(stderr, "  %s %.3f ms\n", what, std::chrono::duration<double, std::milli>(t - T0).count()); T0 = t; } };
    setup_builder(b, c);
    lap("setup");
    for (unsigned i = 0; i < b.A + b.L; i++) {
        if (!advice_out[i]) return ZKW_ERR_INVALID;
        b.adv.push_back(advice_out[i]);
    }
    bool ok = false;
    const U256 vx = load_le(pubkey_x), vy = load_le(pubkey_y), vr = load_le(r), vs = load_le(s), vm = load_le(msg_hash);
    const bool have_traj = b.precompute_denominators(vx, vy, vr, vs, vm);
    lap("denominators");
    const unsigned nthreads = have_traj ? synth_threads() : 1;
    if (nthreads <= 1 || c->st.checkpoints.empty()) {
        b.run(vx, vy, vr, vs, vm, &ok);
        if (b.overflow) return ZKW_ERR_UNSUPPORTED;
        for (unsigned i = 0; i < b.A; i++)
            if (b.rows[i] != c->st.rows[i]) return ZKW_ERR_STATE;     // the layout is data-independent by construction
    } else {
        // prologue here, then the window segments on `nthreads` threads: each starts from the row counters and accumulator
        // cells recorded by the structure pass and from the accumulator VALUES of the projective pre-pass
        Builder::RunCtx rc;
        Builder::EPt acc, facc;
        b.prologue(vx, vy, vr, vs, vm, rc, acc, facc);
        lap("prologue");
        const unsigned nseg = b.num_segments(rc.nw), nw = rc.nw;
        const auto& cps = c->st.checkpoints;
        if (cps.size() != nseg) return ZKW_ERR_STATE;
        for (unsigned i = 0; i < b.A; i++)
            if (b.rows[i] != cps[1].rows[i]) return ZKW_ERR_STATE;
        uint64_t total = 0;
        for (unsigned i = 0; i < b.A; i++) total += c->st.rows[i];
        const uint64_t work = total - cps[1].cells_before;
        // the window segments are cut into chunks of equal cell counts, four per thread, handed out by an atomic counter:
        // the threads do not run at the same speed (hyper-thread siblings, a busy host), and a static split waited for the
        // slowest (per-thread CPU times 0.41-0.81 ms for equal cell counts on the GPU box)
        const unsigned nchunks = std::max(1u, std::min(nseg - 1, 4 * nthreads));
        std::vector<unsigned> cut(nchunks + 1, nseg);
        cut[0] = 1;
        for (unsigned t = 1, sgm = 1; t < nchunks; t++) {
            const uint64_t target = cps[1].cells_before + work * t / nchunks;
            while (sgm < nseg && cps[sgm].cells_before < target) sgm++;
            cut[t] = sgm;
        }
        std::vector<Builder> workers(nthreads, b);        // copies: own cursor, shared output columns and pre-pass results
        std::vector<char> done_ok(nthreads, 0);
        std::atomic<unsigned> next_chunk{0};
        // every worker bumps rows[column] at the end of each region; the copies' tiny heap arrays would sit side by side in
        // one cache line (false sharing), so each gets a line of its own
        for (auto& w : workers) w.rows.reserve(32);
        auto body = [&](unsigned t) {
            Builder& w = workers[t];
            timespec ts0, ts1;
            clock_gettime(CLOCK_THREAD_CPUTIME_ID, &ts0);
            unsigned taken = 0;
            for (;;) {
            const unsigned ci = next_chunk.fetch_add(1);
            if (ci >= nchunks) break;
            const unsigned s0 = cut[ci], s1 = cut[ci + 1];
            if (s0 >= s1) continue;
            taken++;
            const auto& cp = cps[s0];
            w.rows = cp.rows;
            w.lk_cursor = cp.lookups_before;
            w.dinv_next = Builder::first_op_of(s0, nw);
            Builder::EPt a2 = acc, f2 = facc;
            if (s0 > 1) {
                // accumulators at the start of segment s0: variable part after window min(s0, nw) - 1, fixed part after
                // window s0 - nw - 1 (or its constant start point)
                const unsigned va = (s0 < nw ? s0 : nw) - 1;
                a2.x = w.elem_at(cp.acc_ids, w.traj_acc_x[va]);
                a2.y = w.elem_at(cp.acc_ids + 4, w.traj_acc_y[va]);
                if (s0 > nw) {
                    f2.x = w.elem_at(cp.facc_ids, w.traj_facc_x[s0 - nw - 1]);
                    f2.y = w.elem_at(cp.facc_ids + 4, w.traj_facc_y[s0 - nw - 1]);
                }
            }
            bool okt = true;
            w.run_segments(rc, s0, s1, a2, f2, &okt);
            // hand the rows this chunk wrote (the first chunk: the prologue's too) to memory before the DMA reads them
            for (unsigned i = 0; i < w.A; i++) {
                const uint64_t r0 = s0 == 1 ? 0 : cp.rows[i], r1 = w.rows[i];
                if (r1 > r0) writeback_lines(w.adv[i] + 4 * r0, (size_t)(r1 - r0) * 32);
            }
            if (w.L && !w.selector_mode) {
                const uint64_t c0 = s0 == 1 ? 0 : cp.lookups_before, c1 = w.lk_cursor;
                if (c1 > c0)
                    for (unsigned j = 0; j < w.L; j++) {
                        const uint64_t r0 = c0 / w.L, r1 = std::min<uint64_t>((c1 - 1) / w.L + 1, w.u);
                        if (r1 > r0) writeback_lines(w.adv[w.A + j] + 4 * r0, (size_t)(r1 - r0) * 32);
                    }
            }
            if (s1 == nseg) done_ok[t] = okt ? 1 : 2;
            }
            clock_gettime(CLOCK_THREAD_CPUTIME_ID, &ts1);
            if (timing) fprintf(stderr, "    thread %u: %u chunks, cpu %.3f ms\n", t, taken, (ts1.tv_sec - ts0.tv_sec) * 1e3 + (ts1.tv_nsec - ts0.tv_nsec) * 1e-6);
        };
        std::vector<std::thread> threads;
        for (unsigned t = 1; t < nthreads; t++) threads.emplace_back(body, t);
        body(0);
        for (auto& th : threads) th.join();
        lap("segments");
        for (unsigned t = 0; t < nthreads; t++) {
            if (workers[t].overflow) return ZKW_ERR_UNSUPPORTED;
            if (done_ok[t]) ok = done_ok[t] == 1;
            // a worker whose pre-computed inverses stopped matching (cannot happen for a consistent trajectory) fell back to
            // inverting one by one, which is still correct
        }
        b.rows.assign(c->st.rows.begin(), c->st.rows.begin() + b.A);
    }
    lap("lookup copy");
    if (rows_out)
        for (size_t i = 0; i < c->st.rows.size(); i++) rows_out[i] = (size_t)c->st.rows[i];
    if (signature_ok) *signature_ok = ok ? 1 : 0;
    return ZKW_OK;
}
