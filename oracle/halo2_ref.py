"""Python restatement of halo2's create_proof / verify_proof (PSE halo2_proofs v2023_01_20, KZG + GWC)
for the FlexGate + Range constraint system of ECDSACircuit::configure
(halo2-circuits/src/ecc/ecdsa_p256.rs:94-115), with both transcripts the reference uses:
EvmTranscript (keccak, `generate_proof_evm`, ecdsa_p256.rs:365-375) and Blake2bWrite/Challenge255
(`generate_proof`, ecdsa_p256.rs:415-425).  TEST INFRASTRUCTURE ONLY.

What pins this file to the reference:
  * the proof layout (which commitments / evaluations, in which order), the transcript framing, the
    constraint list with its y-ordering, the rotation sets with their v/u ordering and the final
    pairing equation are exactly what proving-server/P256Verifier.yul checks; `verify_proof` below
    ACCEPTS the reference's golden proof (contracts/test/P256Account.t.sol:120) with the VK constants
    of that Yul (tests/test_golden_proof.py), and so does the Yul itself under oracle/yul_evm.py;
  * proofs made by `create_proof` below are accepted by the same `verify_proof`.
Not pinned (no Rust toolchain, no upstream source in /root/reference): the VK digest (upstream hashes
the Rust Debug string of the pinned VK; here a self-defined digest is used, see `vk_digest`), the
blinding RNG (upstream: OsRng), and — for the Blake2b flavour — the compressed point encoding, which is
restated from knowledge of halo2curves 0.3 (x little-endian, bit 255 = parity of y).

Polynomials are lists of canonical python ints.  MSMs go through the C oracle for speed.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field

import numpy as np

from . import cpu
from .keccak import keccak256
from .pyref import FR_DELTA, P, R, EvaluationDomain, g1_add, g1_mul, g1_neg, poly_eval, G1_GEN


# ---------------------------------------------------------------------------------------------------
# constraint-system shape
# ---------------------------------------------------------------------------------------------------
@dataclass
class Shape:
    k: int
    num_advice: int            # gate advice columns A
    num_lookup_advice: int     # dedicated lookup advice columns L (0 = selector mode: q_lookup * a_0)
    num_fixed: int = 1         # constant columns F
    blinding_factors: int = 6

    @property
    def selector_mode(self):
        return self.num_lookup_advice == 0

    @property
    def degree(self):
        return 5 if self.selector_mode else 4

    @property
    def n(self):
        return 1 << self.k

    @property
    def usable_rows(self):
        return self.n - (self.blinding_factors + 1)

    @property
    def last_rotation(self):
        return -(self.blinding_factors + 1)

    @property
    def num_advice_cols(self):
        return self.num_advice + self.num_lookup_advice

    # fixed column indices: constants, table, gate selectors, (q_lookup)
    @property
    def table_col(self):
        return self.num_fixed

    def q_enable_col(self, c):
        return self.num_fixed + 1 + c

    @property
    def q_lookup_col(self):
        return self.num_fixed + 1 + self.num_advice

    @property
    def num_fixed_cols(self):
        return self.num_fixed + 1 + self.num_advice + (1 if self.selector_mode else 0)

    def advice_queries(self):
        q = [(c, r) for c in range(self.num_advice) for r in range(4)]
        q += [(self.num_advice + l, 0) for l in range(self.num_lookup_advice)]
        return q

    def fixed_queries(self):
        return [(c, 0) for c in range(self.num_fixed_cols)]

    def perm_columns(self):
        return [("fixed", c) for c in range(self.num_fixed)] + [("advice", c) for c in range(self.num_advice_cols)]

    @property
    def chunk_len(self):
        return self.degree - 2

    @property
    def num_perm_sets(self):
        return (len(self.perm_columns()) + self.chunk_len - 1) // self.chunk_len

    @property
    def num_lookups(self):
        return self.num_lookup_advice or 1

    @property
    def quotient_pieces(self):
        return self.degree - 1

    def domain(self):
        return EvaluationDomain(self.degree, self.k)


# ---------------------------------------------------------------------------------------------------
# deterministic blinding stream (upstream: OsRng).  element = 254 random bits mod r from splitmix64
# keyed by (seed, stream, index); mirrored by the device prover so proofs can be compared bit for bit
# ---------------------------------------------------------------------------------------------------
_M32 = (1 << 32) - 1


def _rotl32(x, c):
    return ((x << c) | (x >> (32 - c))) & _M32


def _chacha20_block(key_words, stream: int, counter: int):
    """ChaCha20 block function: 32-byte key (8 words), 64-bit block counter (words 12-13), 64-bit nonce = stream id
    (words 14-15) — the original 64/64 layout."""
    s = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key_words) + [counter & _M32, (counter >> 32) & _M32, stream & _M32, (stream >> 32) & _M32]
    x = list(s)

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & _M32; x[d] = _rotl32(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & _M32; x[b] = _rotl32(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & _M32; x[d] = _rotl32(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & _M32; x[b] = _rotl32(x[b] ^ x[c], 7)

    for _ in range(10):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(x[i] + s[i]) & _M32 for i in range(16)]


def seed_words(seed) -> list:
    """An int seed (< 2^64, the tests' deterministic streams) is the 32-byte key padded with zeros; bytes are the key."""
    if isinstance(seed, (bytes, bytearray)):
        assert len(seed) == 32
        return [int.from_bytes(seed[4 * i: 4 * i + 4], "little") for i in range(8)]
    return [seed & _M32, (seed >> 32) & _M32, 0, 0, 0, 0, 0, 0]


def rand_fr(seed, stream: int, index: int) -> int:
    """Blinding scalar `index` of column stream `stream`: one ChaCha20 block (512 bits, little-endian) modulo r."""
    w = _chacha20_block(seed_words(seed), stream, index)
    return sum(v << (32 * i) for i, v in enumerate(w)) % R


STREAM_ADVICE = 1          # + column
STREAM_LOOKUP_A = 1000     # + 2*lookup
STREAM_LOOKUP_S = 1001     # + 2*lookup
STREAM_PERM_Z = 2000       # + set
STREAM_LOOKUP_Z = 3000     # + lookup
STREAM_RANDOM_POLY = 4000


# ---------------------------------------------------------------------------------------------------
# point helpers
# ---------------------------------------------------------------------------------------------------
def commit(srs_points: np.ndarray, values) -> tuple | None:
    """KZG commitment sum_i values[i] * srs[i] (commit over g, commit_lagrange over g_lagrange)."""
    sc = cpu.fr_to_mont(list(values))
    xyz = cpu.best_multiexp(sc, srs_points[: len(values)])
    return cpu.g1_affine_to_ints(cpu.g1_to_affine(xyz)[0])


def g1_compress(pt) -> bytes:
    """halo2curves bn256 G1 compressed form [UPSTREAM, from knowledge]: x little-endian, bit 255 = y & 1."""
    if pt is None:
        return b"\0" * 32
    b = bytearray(pt[0].to_bytes(32, "little"))
    b[31] |= (pt[1] & 1) << 7
    return bytes(b)


def g1_decompress(b: bytes):
    if b == b"\0" * 32:
        return None
    sign = b[31] >> 7
    x = int.from_bytes(b[:31] + bytes([b[31] & 0x7F]), "little")
    y2 = (x * x * x + 3) % P
    y = pow(y2, (P + 1) // 4, P)
    if y * y % P != y2:
        raise ValueError("not on curve")
    if (y & 1) != sign:
        y = P - y
    return (x, y)


# ---------------------------------------------------------------------------------------------------
# transcripts
# ---------------------------------------------------------------------------------------------------
class EvmTranscript:
    """snark-verifier's EvmTranscript<G1Affine, NativeLoader>: keccak256 over a running buffer; points as
    64 bytes (x, y big-endian), scalars as 32 bytes big-endian; after a squeeze the buffer is the 32-byte
    hash, and a squeeze of a bare 32-byte buffer appends the byte 0x01 first (yul:75,97,103-104)."""
    kind = "evm"
    point_bytes = 64

    def __init__(self, proof: bytes | None = None):
        self.buf = bytearray()
        self.out = bytearray()
        self.inp = proof
        self.pos = 0

    def common_scalar(self, s: int):
        self.buf += s.to_bytes(32, "big")

    def common_point(self, pt):
        assert pt is not None
        self.buf += pt[0].to_bytes(32, "big") + pt[1].to_bytes(32, "big")

    def write_point(self, pt):
        self.common_point(pt)
        self.out += pt[0].to_bytes(32, "big") + pt[1].to_bytes(32, "big")

    def write_scalar(self, s: int):
        self.common_scalar(s)
        self.out += s.to_bytes(32, "big")

    def read_point(self):
        x = int.from_bytes(self.inp[self.pos:self.pos + 32], "big")
        y = int.from_bytes(self.inp[self.pos + 32:self.pos + 64], "big")
        self.pos += 64
        if x >= P or y >= P or (y * y - x * x * x - 3) % P:
            raise ValueError("invalid point in proof")
        self.common_point((x, y))
        return (x, y)

    def read_scalar(self):
        s = int.from_bytes(self.inp[self.pos:self.pos + 32], "big")
        self.pos += 32
        if s >= R:
            raise ValueError("non-canonical scalar in proof")
        self.common_scalar(s)
        return s

    def squeeze(self) -> int:
        data = bytes(self.buf) + (b"\x01" if len(self.buf) == 32 else b"")
        h = keccak256(data)
        self.buf = bytearray(h)
        return int.from_bytes(h, "big") % R


class Blake2bTranscript:
    """halo2_proofs::transcript::Blake2bWrite / Blake2bRead with Challenge255: one running
    Blake2b-512 state personalised "Halo2-Transcript"; prefix bytes 0 (challenge), 1 (point), 2 (scalar);
    coordinates and scalars little-endian; challenge = 64-byte digest mod r; points are written
    compressed (32 bytes)."""
    kind = "blake2b"
    point_bytes = 32

    def __init__(self, proof: bytes | None = None):
        self.state = hashlib.blake2b(digest_size=64, person=b"Halo2-Transcript")
        self.out = bytearray()
        self.inp = proof
        self.pos = 0

    def common_scalar(self, s: int):
        self.state.update(b"\x02" + s.to_bytes(32, "little"))

    def common_point(self, pt):
        assert pt is not None
        self.state.update(b"\x01" + pt[0].to_bytes(32, "little") + pt[1].to_bytes(32, "little"))

    def write_point(self, pt):
        self.common_point(pt)
        self.out += g1_compress(pt)

    def write_scalar(self, s: int):
        self.common_scalar(s)
        self.out += s.to_bytes(32, "little")

    def read_point(self):
        pt = g1_decompress(self.inp[self.pos:self.pos + 32])
        self.pos += 32
        self.common_point(pt)
        return pt

    def read_scalar(self):
        s = int.from_bytes(self.inp[self.pos:self.pos + 32], "little")
        self.pos += 32
        if s >= R:
            raise ValueError("non-canonical scalar in proof")
        self.common_scalar(s)
        return s

    def squeeze(self) -> int:
        self.state.update(b"\x00")
        h = self.state.copy().digest()
        return int.from_bytes(h, "little") % R


def make_transcript(kind: str, proof: bytes | None = None):
    return EvmTranscript(proof) if kind == "evm" else Blake2bTranscript(proof)


# ---------------------------------------------------------------------------------------------------
# keys
# ---------------------------------------------------------------------------------------------------
@dataclass
class VerifyingKey:
    shape: Shape
    fixed_commitments: list
    perm_commitments: list
    digest: int
    g0: tuple = G1_GEN


@dataclass
class ProvingKey:
    vk: VerifyingKey
    fixed_values: list       # lagrange
    fixed_polys: list        # coefficients
    sigma_values: list       # lagrange values delta^c' * omega^r'
    sigma_polys: list
    l0: list = field(default_factory=list)
    l_last: list = field(default_factory=list)
    l_active: list = field(default_factory=list)


def vk_digest(shape: Shape, fixed_commitments, perm_commitments) -> int:
    """Self-defined transcript_repr (upstream: Blake2b("Halo2-Verify-Key") over the Debug string of the
    pinned VK, which cannot be reproduced outside Rust): Blake2b-512 with the same personalisation over a
    canonical binary serialisation, reduced mod r."""
    h = hashlib.blake2b(digest_size=64, person=b"Halo2-Verify-Key")
    h.update(b"zkw-b200-vk-v1")
    for v in (shape.k, shape.num_advice, shape.num_lookup_advice, shape.num_fixed, shape.blinding_factors, shape.degree):
        h.update(int(v).to_bytes(4, "little"))
    for pt in list(fixed_commitments) + list(perm_commitments):
        x, y = pt if pt is not None else (0, 0)
        h.update(x.to_bytes(32, "little") + y.to_bytes(32, "little"))
    return int.from_bytes(h.digest(), "little") % R


def identity_sigma(shape: Shape):
    """sigma with no copy constraints: sigma_c(row) = delta^c * omega^row."""
    dom = shape.domain()
    out = []
    for c in range(len(shape.perm_columns())):
        d = pow(FR_DELTA, c, R)
        w, col = 1, []
        for _ in range(shape.n):
            col.append(d * w % R)
            w = w * dom.omega % R
        out.append(col)
    return out


def sigma_from_cycles(shape: Shape, mapping):
    """mapping[c][row] = (c', row'): the permutation as halo2's keygen Assembly stores it."""
    dom = shape.domain()
    pw = [1] * shape.n
    for i in range(1, shape.n):
        pw[i] = pw[i - 1] * dom.omega % R
    dl = [pow(FR_DELTA, c, R) for c in range(len(mapping))]
    return [[dl[cc] * pw[rr] % R for (cc, rr) in col] for col in mapping]


def keygen(shape: Shape, g_lagrange: np.ndarray, fixed_values, sigma_values) -> ProvingKey:
    dom = shape.domain()
    n = shape.n
    assert len(fixed_values) == shape.num_fixed_cols and len(sigma_values) == len(shape.perm_columns())
    fixed_commitments = [commit(g_lagrange, col) for col in fixed_values]
    perm_commitments = [commit(g_lagrange, col) for col in sigma_values]
    vk = VerifyingKey(shape, fixed_commitments, perm_commitments, vk_digest(shape, fixed_commitments, perm_commitments))
    l0 = [0] * n
    l0[0] = 1
    l_last = [0] * n
    l_last[shape.usable_rows] = 1
    l_active = [1 if i < shape.usable_rows else 0 for i in range(n)]
    return ProvingKey(vk, fixed_values, [dom.lagrange_to_coeff(c) for c in fixed_values], sigma_values,
                      [dom.lagrange_to_coeff(c) for c in sigma_values], l0, l_last, l_active)


# ---------------------------------------------------------------------------------------------------
# prover
# ---------------------------------------------------------------------------------------------------
def lookup_input_values(shape: Shape, advice, fixed, l: int):
    if shape.selector_mode:
        q = fixed[shape.q_lookup_col]
        return [q[i] * advice[0][i] % R for i in range(shape.n)]
    return list(advice[shape.num_advice + l])


def permute_expression_pair(shape: Shape, inp, table, seed, l):
    """halo2_proofs::plonk::lookup::prover::permute_expression_pair restated: A' = sorted input over the
    usable rows; S' holds the table value at the first row of every run of A' and the leftover table
    values (ascending) at the remaining rows taken from the LAST one backwards; then blinding rows."""
    u = shape.usable_rows
    a = sorted(inp[:u])
    left = {}
    for t in table[:u]:
        left[t] = left.get(t, 0) + 1
    s = [0] * u
    repeated = []
    for row in range(u):
        if row == 0 or a[row] != a[row - 1]:
            s[row] = a[row]
            if left.get(a[row], 0) <= 0:
                raise ValueError("lookup input not in table (ConstraintSystemFailure)")
            left[a[row]] -= 1
        else:
            repeated.append(row)
    for val in sorted(left):
        for _ in range(left[val]):
            s[repeated.pop()] = val
    assert not repeated
    nb = shape.blinding_factors + 1
    a += [rand_fr(seed, STREAM_LOOKUP_A + 2 * l, i) for i in range(nb)]
    s += [rand_fr(seed, STREAM_LOOKUP_S + 2 * l, i) for i in range(nb)]
    return a, s


def batch_inv(vals):
    out = [0] * len(vals)
    acc = 1
    pref = []
    for v in vals:
        pref.append(acc)
        if v:
            acc = acc * v % R
    inv = pow(acc, -1, R)
    for i in range(len(vals) - 1, -1, -1):
        if vals[i]:
            out[i] = inv * pref[i] % R
            inv = inv * vals[i] % R
    return out


def kate_division(coeffs, z):
    """(p(X) - p(z)) / (X - z) by synthetic division (halo2_proofs::arithmetic::kate_division)."""
    q = [0] * (len(coeffs) - 1)
    carry = 0
    for i in range(len(coeffs) - 1, 0, -1):
        carry = (coeffs[i] + carry * z) % R
        q[i - 1] = carry
    return q


def lagrange_interpolate(points, evals):
    """coefficients of the unique polynomial of degree < len(points) through (points[i], evals[i])."""
    m = len(points)
    out = [0] * m
    for i in range(m):
        num = [1]
        den = 1
        for j in range(m):
            if j == i:
                continue
            num = [(a - points[j] * b) % R for a, b in zip([0] + num, num + [0])]   # num * (X - p_j)
            den = den * (points[i] - points[j]) % R
        c = evals[i] * pow(den, -1, R) % R
        for d in range(len(num)):
            out[d] = (out[d] + c * num[d]) % R
    return out


def _shplonk_sets(queries, point_of):
    """halo2_proofs::poly::kzg::multiopen::shplonk::construct_intermediate_sets restated.
    queries: (rot, key, payload).  Returns (sets, super_points): sets in order of first appearance of
    each distinct point set; inside a set the points are in ascending field order (BTreeSet) and the
    commitments in order of first appearance."""
    comm = []
    super_points = set()
    for rot, key, payload in queries:
        pt = point_of(rot)
        super_points.add(pt)
        for e in comm:
            if e[0] == key:
                e[2].add(pt)
                break
        else:
            comm.append((key, payload, {pt}))
    sets = []
    for key, payload, pts in comm:
        sk = tuple(sorted(pts))
        for st in sets:
            if st[0] == sk:
                st[1].append((key, payload))
                break
        else:
            sets.append((sk, [(key, payload)]))
    return sets, sorted(super_points)


def shplonk_prove(tr, g, queries, point_of, n):
    """ProverSHPLONK::create_proof restated.  queries: (rot, poly) with poly identity = list identity."""
    y = tr.squeeze()
    v = tr.squeeze()
    sets, super_points = _shplonk_sets([(r, id(p), p) for r, p in queries], point_of)
    interp = []
    h_x = [0] * n
    pv = 1
    for points, polys in sets:
        n_x = [0] * n
        py = 1
        rs = []
        for _, poly in polys:
            r_x = lagrange_interpolate(list(points), [poly_eval(poly, pt) for pt in points])
            rs.append(r_x)
            for i in range(n):
                n_x[i] = (n_x[i] + py * (poly[i] - (r_x[i] if i < len(r_x) else 0))) % R
            py = py * y % R
        interp.append(rs)
        q = n_x
        for pt in points:
            q = kate_division(q, pt)
        q = q + [0] * (n - len(q))
        h_x = [(a + pv * b) % R for a, b in zip(h_x, q)]
        pv = pv * v % R
    tr.write_point(commit(g, h_x))
    u = tr.squeeze()
    l_x = [0] * n
    z_diffs = []
    pv = 1
    for (points, polys), rs in zip(sets, interp):
        z_i = 1
        for d in super_points:
            if d not in points:
                z_i = z_i * (u - d) % R
        z_diffs.append(z_i)
        py = 1
        acc = [0] * n
        for (_, poly), r_x in zip(polys, rs):
            r_eval = poly_eval(r_x, u)
            for i in range(n):
                acc[i] = (acc[i] + py * poly[i]) % R
            acc[0] = (acc[0] - py * r_eval) % R
            py = py * y % R
        l_x = [(a + pv * z_i % R * b) % R for a, b in zip(l_x, acc)]
        pv = pv * v % R
    zt = 1
    for d in super_points:
        zt = zt * (u - d) % R
    l_x = [(a - zt * b) % R for a, b in zip(l_x, h_x)]
    assert poly_eval(l_x, u) == 0
    hq = kate_division(l_x, u)
    inv0 = pow(z_diffs[0], -1, R)
    tr.write_point(commit(g, [c * inv0 % R for c in hq]))


def create_proof(pk: ProvingKey, g: np.ndarray, g_lagrange: np.ndarray, advice_in, seed: int, kind: str, multiopen: str = "gwc") -> bytes:
    """advice_in: the advice columns' usable rows (lists of ints, length <= usable_rows each).
    multiopen: "gwc" (ProverGWC, what generate_proof_evm uses) or "shplonk" (ProverSHPLONK, generate_proof)."""
    shape = pk.vk.shape
    dom = shape.domain()
    n, u = shape.n, shape.usable_rows
    omega = dom.omega
    tr = make_transcript(kind)
    tr.common_scalar(pk.vk.digest)

    # 1. advice columns: witness rows, zero padding, blinding rows, commitments
    advice = []
    for c in range(shape.num_advice_cols):
        col = list(advice_in[c]) + [0] * (u - len(advice_in[c]))
        col += [rand_fr(seed, STREAM_ADVICE + c, i) for i in range(n - u)]
        advice.append(col)
    for col in advice:
        tr.write_point(commit(g_lagrange, col))
    theta = tr.squeeze()  # noqa: F841  (single-expression lookups: compression is the identity)

    # 2. lookups: permuted input / table
    lookups = []
    for l in range(shape.num_lookups):
        inp = lookup_input_values(shape, advice, pk.fixed_values, l)
        tab = pk.fixed_values[shape.table_col]
        a, s = permute_expression_pair(shape, inp, tab, seed, l)
        lookups.append({"inp": inp, "tab": tab, "a": a, "s": s})
        tr.write_point(commit(g_lagrange, a))
        tr.write_point(commit(g_lagrange, s))
    beta = tr.squeeze()
    gamma = tr.squeeze()

    # 3. permutation grand products
    cols = shape.perm_columns()
    values = [pk.fixed_values[c] if t == "fixed" else advice[c] for t, c in cols]
    pw = [1] * n
    for i in range(1, n):
        pw[i] = pw[i - 1] * omega % R
    perm_z = []
    last_z = 1
    for s_idx in range(shape.num_perm_sets):
        c0, c1 = s_idx * shape.chunk_len, min((s_idx + 1) * shape.chunk_len, len(cols))
        den = [1] * n
        num = [1] * n
        for c in range(c0, c1):
            dc = pow(FR_DELTA, c, R) * beta % R
            for i in range(n):
                den[i] = den[i] * (values[c][i] + beta * pk.sigma_values[c][i] + gamma) % R
                num[i] = num[i] * (values[c][i] + dc * pw[i] + gamma) % R
        den = batch_inv(den)
        z = [last_z]
        for i in range(n - 1):
            z.append(z[-1] * num[i] % R * den[i] % R)
        for i in range(u + 1, n):
            z[i] = rand_fr(seed, STREAM_PERM_Z + s_idx, i - (u + 1))
        last_z = z[u]
        perm_z.append(z)
    # 4. lookup grand products
    for l, lk in enumerate(lookups):
        den = batch_inv([(lk["a"][i] + beta) * (lk["s"][i] + gamma) % R for i in range(n)])
        z = [1]
        for i in range(n - 1):
            z.append(z[-1] * ((lk["inp"][i] + beta) * (lk["tab"][i] + gamma) % R) % R * den[i] % R)
        for i in range(u + 1, n):
            z[i] = rand_fr(seed, STREAM_LOOKUP_Z + l, i - (u + 1))
        lk["z"] = z
    for z in perm_z:
        tr.write_point(commit(g_lagrange, z))
    for lk in lookups:
        tr.write_point(commit(g_lagrange, lk["z"]))

    # 5. vanishing argument: random polynomial
    random_poly = [rand_fr(seed, STREAM_RANDOM_POLY, i) for i in range(n)]
    tr.write_point(commit(g, random_poly))
    y = tr.squeeze()

    # 6. quotient
    to_coeff = dom.lagrange_to_coeff
    advice_polys = [to_coeff(c) for c in advice]
    perm_z_polys = [to_coeff(z) for z in perm_z]
    for lk in lookups:
        lk["z_poly"], lk["a_poly"], lk["s_poly"] = to_coeff(lk["z"]), to_coeff(lk["a"]), to_coeff(lk["s"])
    ext = dom.coeff_to_extended
    from .pyref import quotient_ecdsa
    qcols = {
        "advice": [ext(p) for p in advice_polys],
        "constants": [ext(pk.fixed_polys[c]) for c in range(shape.num_fixed)],
        "table": ext(pk.fixed_polys[shape.table_col]),
        "q_enable": [ext(pk.fixed_polys[shape.q_enable_col(c)]) for c in range(shape.num_advice)],
        "q_lookup": ext(pk.fixed_polys[shape.q_lookup_col]) if shape.selector_mode else None,
        "sigma": [ext(p) for p in pk.sigma_polys],
        "perm_z": [ext(p) for p in perm_z_polys],
        "lookup_z": [ext(lk["z_poly"]) for lk in lookups],
        "lookup_a": [ext(lk["a_poly"]) for lk in lookups],
        "lookup_s": [ext(lk["s_poly"]) for lk in lookups],
        "l0": ext(to_coeff(pk.l0)), "l_last": ext(to_coeff(pk.l_last)), "l_active": ext(to_coeff(pk.l_active)),
    }
    pshape = {"k": shape.k, "ext_k": dom.extended_k, "num_advice": shape.num_advice, "num_lookup_advice": shape.num_lookup_advice,
              "num_fixed": shape.num_fixed, "blinding_factors": shape.blinding_factors, "cs_degree": shape.degree}
    h_ext = quotient_ecdsa(pshape, qcols, {"y": y, "beta": beta, "gamma": gamma})
    h_coeff = dom.extended_to_coeff(h_ext)
    assert len(h_coeff) == n * shape.quotient_pieces
    h_pieces = [h_coeff[i * n:(i + 1) * n] for i in range(shape.quotient_pieces)]
    for piece in h_pieces:
        tr.write_point(commit(g, piece))
    x = tr.squeeze()

    # 7. evaluations
    rot = lambda r: x * pow(omega, r, R) % R
    for c, r in shape.advice_queries():
        tr.write_scalar(poly_eval(advice_polys[c], rot(r)))
    for c, r in shape.fixed_queries():
        tr.write_scalar(poly_eval(pk.fixed_polys[c], rot(r)))
    xn = pow(x, n, R)
    h_poly = [0] * n
    for piece in reversed(h_pieces):
        h_poly = [(a * xn + b) % R for a, b in zip(h_poly, piece)]
    tr.write_scalar(poly_eval(random_poly, x))
    for p in pk.sigma_polys:
        tr.write_scalar(poly_eval(p, x))
    for s_idx, zp in enumerate(perm_z_polys):
        tr.write_scalar(poly_eval(zp, x))
        tr.write_scalar(poly_eval(zp, rot(1)))
        if s_idx != len(perm_z_polys) - 1:
            tr.write_scalar(poly_eval(zp, rot(shape.last_rotation)))
    for lk in lookups:
        tr.write_scalar(poly_eval(lk["z_poly"], x))
        tr.write_scalar(poly_eval(lk["z_poly"], rot(1)))
        tr.write_scalar(poly_eval(lk["a_poly"], x))
        tr.write_scalar(poly_eval(lk["a_poly"], rot(-1)))
        tr.write_scalar(poly_eval(lk["s_poly"], x))

    # 8. multi-open (GWC): queries in upstream order, grouped by point in order of first appearance
    queries = []
    for c, r in shape.advice_queries():
        queries.append((r, advice_polys[c]))
    for zp in perm_z_polys:
        queries.append((0, zp))
        queries.append((1, zp))
    for zp in list(reversed(perm_z_polys))[1:]:
        queries.append((shape.last_rotation, zp))
    for lk in lookups:
        queries += [(0, lk["z_poly"]), (0, lk["a_poly"]), (0, lk["s_poly"]), (-1, lk["a_poly"]), (1, lk["z_poly"])]
    for c, r in shape.fixed_queries():
        queries.append((r, pk.fixed_polys[c]))
    for p in pk.sigma_polys:
        queries.append((0, p))
    queries.append((0, h_poly))
    queries.append((0, random_poly))
    if multiopen == "shplonk":
        shplonk_prove(tr, g, queries, rot, n)
        return bytes(tr.out)
    v = tr.squeeze()
    order = []
    for r, _ in queries:
        if r not in order:
            order.append(r)
    for r in order:
        z = rot(r)
        acc = [0] * n
        pv = 1
        for rr, poly in queries:
            if rr != r:
                continue
            acc = [(a + pv * b) % R for a, b in zip(acc, poly)]
            pv = pv * v % R
        # poly_batch - eval_batch then divide: synthetic division drops the remainder anyway
        tr.write_point(commit(g, kate_division(acc, z)))
    return bytes(tr.out)


# ---------------------------------------------------------------------------------------------------
# verifier
# ---------------------------------------------------------------------------------------------------
def verify_proof(vk: VerifyingKey, proof: bytes, kind: str, *, tau: int | None = None, g2_pair=None, multiopen: str = "gwc") -> bool:
    """halo2 verify_proof + VerifierGWC restated.  The final check is either the pairing
    e(left, [s]_2) == e(right, [1]_2) with g2_pair = (G2, sG2), or — with a development SRS whose tau is
    known — the equivalent G1 identity tau * left == right."""
    shape = vk.shape
    dom = shape.domain()
    n = shape.n
    omega = dom.omega
    try:
        tr = make_transcript(kind, proof)
        tr.common_scalar(vk.digest)
        advice_c = [tr.read_point() for _ in range(shape.num_advice_cols)]
        theta = tr.squeeze()  # noqa: F841
        lk_c = [{"a": tr.read_point(), "s": tr.read_point()} for _ in range(shape.num_lookups)]
        beta = tr.squeeze()
        gamma = tr.squeeze()
        perm_z_c = [tr.read_point() for _ in range(shape.num_perm_sets)]
        for lk in lk_c:
            lk["z"] = tr.read_point()
        random_c = tr.read_point()
        y = tr.squeeze()
        h_c = [tr.read_point() for _ in range(shape.quotient_pieces)]
        x = tr.squeeze()
        advice_e = {q: tr.read_scalar() for q in shape.advice_queries()}
        fixed_e = {q[0]: tr.read_scalar() for q in shape.fixed_queries()}
        random_e = tr.read_scalar()
        sigma_e = [tr.read_scalar() for _ in range(len(shape.perm_columns()))]
        perm_e = []
        for s_idx in range(shape.num_perm_sets):
            e = {"z": tr.read_scalar(), "zn": tr.read_scalar()}
            if s_idx != shape.num_perm_sets - 1:
                e["zl"] = tr.read_scalar()
            perm_e.append(e)
        lk_e = [{"z": tr.read_scalar(), "zn": tr.read_scalar(), "a": tr.read_scalar(), "ai": tr.read_scalar(), "s": tr.read_scalar()}
                for _ in range(shape.num_lookups)]
    except (ValueError, IndexError):
        return False

    # Lagrange evaluations at x
    xn = pow(x, n, R)
    if xn == 1:
        return False
    zh_n = (xn - 1) * pow(n, -1, R) % R

    def lagrange(i):  # l_i(x) for row i (negative = from the end)
        wi = pow(omega, i, R)
        return zh_n * wi % R * pow((x - wi) % R, -1, R) % R

    l0 = lagrange(0)
    l_last = lagrange(shape.last_rotation)
    l_blind = sum(lagrange(-j) for j in range(1, shape.blinding_factors + 1)) % R
    l_active = (1 - l_last - l_blind) % R

    # constraints, folded with y in upstream order
    acc = 0

    def fold(c):
        nonlocal acc
        acc = (acc * y + c) % R

    for c in range(shape.num_advice):
        a = [advice_e[(c, r)] for r in range(4)]
        fold(fixed_e[shape.q_enable_col(c)] * (a[0] + a[1] * a[2] - a[3]))
    cols = shape.perm_columns()
    col_e = [fixed_e[c] if t == "fixed" else advice_e[(c, 0)] for t, c in cols]
    fold(l0 * (1 - perm_e[0]["z"]))
    zl = perm_e[-1]["z"]
    fold(l_last * (zl * zl - zl))
    for s_idx in range(1, shape.num_perm_sets):
        fold(l0 * (perm_e[s_idx]["z"] - perm_e[s_idx - 1]["zl"]))
    for s_idx in range(shape.num_perm_sets):
        c0, c1 = s_idx * shape.chunk_len, min((s_idx + 1) * shape.chunk_len, len(cols))
        left, right = perm_e[s_idx]["zn"], perm_e[s_idx]["z"]
        for c in range(c0, c1):
            left = left * (col_e[c] + beta * sigma_e[c] + gamma) % R
            right = right * (col_e[c] + pow(FR_DELTA, c, R) * beta % R * x + gamma) % R
        fold(l_active * (left - right))
    for l, e in enumerate(lk_e):
        inp = fixed_e[shape.q_lookup_col] * advice_e[(0, 0)] % R if shape.selector_mode else advice_e[(shape.num_advice + l, 0)]
        tab = fixed_e[shape.table_col]
        fold(l0 * (1 - e["z"]))
        fold(l_last * (e["z"] * e["z"] - e["z"]))
        fold(l_active * (e["zn"] * (e["a"] + beta) % R * (e["s"] + gamma) - e["z"] * (inp + beta) % R * (tab + gamma)))
        fold(l0 * (e["a"] - e["s"]))
        fold(l_active * ((e["a"] - e["s"]) * (e["a"] - e["ai"]) % R))
    h_eval = acc * pow((xn - 1) % R, -1, R) % R

    # h commitment: sum_i x^(n i) H_i
    h_commit = None
    for hc in reversed(h_c):
        h_commit = g1_add(g1_mul(h_commit, xn), hc)

    # queries in the prover's order: (rotation, commitment, eval)
    queries = []
    for q in shape.advice_queries():
        queries.append((q[1], advice_c[q[0]], advice_e[q]))
    for s_idx in range(shape.num_perm_sets):
        queries.append((0, perm_z_c[s_idx], perm_e[s_idx]["z"]))
        queries.append((1, perm_z_c[s_idx], perm_e[s_idx]["zn"]))
    for s_idx in reversed(range(shape.num_perm_sets - 1)):
        queries.append((shape.last_rotation, perm_z_c[s_idx], perm_e[s_idx]["zl"]))
    for lk, e in zip(lk_c, lk_e):
        queries += [(0, lk["z"], e["z"]), (0, lk["a"], e["a"]), (0, lk["s"], e["s"]), (-1, lk["a"], e["ai"]), (1, lk["z"], e["zn"])]
    for q in shape.fixed_queries():
        queries.append((q[1], vk.fixed_commitments[q[0]], fixed_e[q[0]]))
    for c, pc in enumerate(vk.perm_commitments):
        queries.append((0, pc, sigma_e[c]))
    queries.append((0, h_commit, h_eval))
    queries.append((0, random_c, random_e))

    if multiopen == "shplonk":
        # VerifierSHPLONK restated; a commitment's identity is its position in the query list's first use
        keyed = []
        seen = {}
        for r, cm, ev in queries:
            k = seen.setdefault(id(cm) if cm is not None else ("none", len(seen)), len(seen))
            keyed.append((r, k, cm, ev))
        point_of = lambda r: x * pow(omega, r, R) % R
        y_ch = tr.squeeze()
        v_ch = tr.squeeze()
        try:
            h1 = tr.read_point()
            u_ch = tr.squeeze()
            h2 = tr.read_point()
        except (ValueError, IndexError):
            return False
        if tr.pos != len(proof):
            return False
        evals_of = {}
        for r, k, cm, ev in keyed:
            evals_of[(k, point_of(r))] = ev
        sets, super_points = _shplonk_sets([(r, k, cm) for r, k, cm, ev in keyed], point_of)
        outer = None
        r_outer = 0
        z_0 = z_0_diff_inv = 0
        pv = 1
        for i, (points, comms) in enumerate(sets):
            z_diff = 1
            for d in super_points:
                if d not in points:
                    z_diff = z_diff * (u_ch - d) % R
            if i == 0:
                z_0 = 1
                for pt in points:
                    z_0 = z_0 * (u_ch - pt) % R
                z_0_diff_inv = pow(z_diff, -1, R)
                z_diff = 1
            else:
                z_diff = z_diff * z_0_diff_inv % R
            inner, r_inner, py = None, 0, 1
            for k, cm in comms:
                r_x = lagrange_interpolate(list(points), [evals_of[(k, pt)] for pt in points])
                r_inner = (r_inner + py * poly_eval(r_x, u_ch)) % R
                inner = g1_add(inner, g1_mul(cm, py))
                py = py * y_ch % R
            outer = g1_add(outer, g1_mul(inner, pv * z_diff % R))
            r_outer = (r_outer + pv * r_inner % R * z_diff) % R
            pv = pv * v_ch % R
        right = g1_add(outer, g1_neg(g1_mul(vk.g0, r_outer)))
        right = g1_add(right, g1_neg(g1_mul(h1, z_0)))
        right = g1_add(right, g1_mul(h2, u_ch))
        left = h2
        if tau is not None:
            return g1_mul(left, tau) == right
        from . import pairing
        g2, s_g2 = g2_pair
        return pairing.pairing_product_is_one([(left, s_g2), (g1_neg(right), g2)])

    v = tr.squeeze()
    order = []
    for r, _, _ in queries:
        if r not in order:
            order.append(r)
    try:
        ws = [tr.read_point() for _ in order]
    except (ValueError, IndexError):
        return False
    if tr.pos != len(proof):
        return False
    u = tr.squeeze()
    left = None          # sum u^i W_i
    right = None         # sum u^i (z_i W_i + C_i) - (sum u^i e_i) G
    eval_multi = 0
    pu = 1
    for r, w in zip(order, ws):
        z = x * pow(omega, r, R) % R
        cb, eb, pv = None, 0, 1
        for rr, cm, ev in queries:
            if rr != r:
                continue
            cb = g1_add(cb, g1_mul(cm, pv))
            eb = (eb + pv * ev) % R
            pv = pv * v % R
        right = g1_add(right, g1_mul(cb, pu))
        eval_multi = (eval_multi + pu * eb) % R
        right = g1_add(right, g1_mul(w, pu * z % R))
        left = g1_add(left, g1_mul(w, pu))
        pu = pu * u % R
    right = g1_add(right, g1_neg(g1_mul(vk.g0, eval_multi)))
    if tau is not None:
        return g1_mul(left, tau) == right
    from . import pairing
    g2, s_g2 = g2_pair
    return pairing.pairing_product_is_one([(left, s_g2), (g1_neg(right), g2)])
