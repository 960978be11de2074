"""Python big-integer restatement of the BN254 arithmetic under halo2's prover hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
path (``webauthn-halo2_b200/``); only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may use it, and
there only as the checker.

PARITY STATUS: *unpinned at the MSM / NTT / quotient boundary* (SURVEY.md §8c).
The reference (`/root/reference`) holds no MSM, NTT or h(X) known-answer: its
hot path lives in un-vendored, un-pinned crates reached through
`halo2-circuits/Cargo.toml:12-15` (halo2-lib@main -> PSE halo2_proofs
v2023_01_20 + halo2curves 0.3.x; snark-verifier@v2023_01_20_secp256r1).  What
the reference *does* pin, and what this file is checked against in
``tests/test_oracle_constants.py``:

  * p, r                         proving-server/P256Verifier.yul:17-18
  * n^-1 for k=17                P256Verifier.yul:307  (0xfa0 multiplier)
  * omega_17^-j, j=0..7          P256Verifier.yul:308-323
  * delta^1..delta^5             P256Verifier.yul:465,483,487,505,509
  * the whole verification algorithm: golden proof + Yul (tests/test_golden_proof.py)

Everything below restates published algorithms (halo2_proofs::arithmetic,
poly::EvaluationDomain, halo2curves::bn256) from their definitions; outputs are
canonical (affine points, field elements), so algorithm choice is free.

Slow and simple on purpose: this file pins the C oracle (oracle/*.c), which in
turn pins the CUDA kernels at sizes Python cannot reach.
"""
from __future__ import annotations

import math

# --- BN254 moduli (P256Verifier.yul:17-18: f_p, f_q) -------------------------------
P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47  # base field Fq
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001  # scalar field Fr

# halo2curves::bn256::Fr constants: MULTIPLICATIVE_GENERATOR = 7, S = 28
FR_GENERATOR = 7
FR_S = 28
FR_ROOT_OF_UNITY = pow(FR_GENERATOR, (R - 1) >> FR_S, R)          # 2^28-th primitive root
FR_DELTA = pow(FR_GENERATOR, 1 << FR_S, R)                         # generator of the t-order subgroup
FR_ZETA = pow(FR_GENERATOR, (R - 1) // 3, R)                       # primitive cube root of unity
assert FR_ZETA == 0xB3C4D79D41A917585BFC41088D8DAAA78B17EA66B99C90DD

MONT_R_BITS = 256
FR_MONT_R = (1 << MONT_R_BITS) % R
FQ_MONT_R = (1 << MONT_R_BITS) % P
FR_INV64 = (-pow(R, -1, 1 << 64)) % (1 << 64)
FQ_INV64 = (-pow(P, -1, 1 << 64)) % (1 << 64)

G1_GEN = (1, 2)           # halo2curves bn256 G1 generator; curve y^2 = x^3 + 3
G1_B = 3


# --- Montgomery (de)serialisation: halo2curves stores [u64;4] LE limbs, Montgomery form ----
def to_mont(x: int, mod: int) -> int:
    return (x << MONT_R_BITS) % mod


def from_mont(x: int, mod: int) -> int:
    return (x * pow(1 << MONT_R_BITS, -1, mod)) % mod


def limbs64(x: int) -> list[int]:
    return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def from_limbs64(l) -> int:
    return sum(int(v) << (64 * i) for i, v in enumerate(l))


# --- G1 arithmetic, affine with None as identity (canonical form) -------------------------
def g1_is_on_curve(pt) -> bool:
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - G1_B) % P == 0


def g1_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return None
        lam = (3 * x1 * x1) * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    y3 = (lam * (x1 - x3) - y1) % P
    return (x3, y3)


def g1_neg(a):
    return None if a is None else (a[0], (-a[1]) % P)


def g1_mul(a, k: int):
    k %= R
    acc = None
    while k:
        if k & 1:
            acc = g1_add(acc, a)
        a = g1_add(a, a)
        k >>= 1
    return acc


def msm_naive(scalars, bases):
    acc = None
    for s, b in zip(scalars, bases):
        acc = g1_add(acc, g1_mul(b, s))
    return acc


def multiexp_serial(scalars, bases, acc):
    """halo2_proofs::arithmetic::multiexp_serial restated (windowed Pippenger).

    Window c = 1 (n<4), 3 (n<32), else ceil(ln n); segments = 256/c + 1;
    2^c - 1 buckets; running-sum reduction from the top bucket down.
    Called from best_multiexp via ParamsKZG::commit{,_lagrange}, which the
    reference reaches through create_proof (halo2-circuits/src/ecc/ecdsa_p256.rs:366,416,555).
    """
    n = len(bases)
    if n < 4:
        c = 1
    elif n < 32:
        c = 3
    else:
        c = math.ceil(math.log(n))
    segments = 256 // c + 1
    for seg in reversed(range(segments)):
        for _ in range(c):
            acc = g1_add(acc, acc)
        buckets = [None] * ((1 << c) - 1)
        for s, b in zip(scalars, bases):
            d = (s >> (seg * c)) & ((1 << c) - 1)
            if d:
                buckets[d - 1] = g1_add(buckets[d - 1], b)
        running = None
        for bk in reversed(buckets):
            running = g1_add(running, bk)
            acc = g1_add(acc, running)
    return acc


def best_multiexp(scalars, bases, threads: int = 1):
    """halo2_proofs::arithmetic::best_multiexp restated: one contiguous chunk per thread."""
    n = len(scalars)
    assert n == len(bases)
    if n > threads:
        chunk = n // threads
        acc = None
        for i in range(0, n, chunk):
            acc = g1_add(acc, multiexp_serial(scalars[i:i + chunk], bases[i:i + chunk], None))
        return acc
    return multiexp_serial(scalars, bases, None)


# --- NTT ------------------------------------------------------------------------------------
def bitreverse(n: int, l: int) -> int:
    r = 0
    for _ in range(l):
        r = (r << 1) | (n & 1)
        n >>= 1
    return r


def best_fft(a, omega: int, log_n: int):
    """halo2_proofs::arithmetic::best_fft contract: natural order in and out,
    out[i] = sum_j a[j] * omega^(i*j).  Bit-reversal permutation followed by
    radix-2 decimation-in-time butterflies (the serial branch of the upstream routine)."""
    n = 1 << log_n
    a = list(a)
    assert len(a) == n
    for k in range(n):
        rk = bitreverse(k, log_n)
        if k < rk:
            a[k], a[rk] = a[rk], a[k]
    tw = [1] * (n // 2)
    for i in range(1, n // 2):
        tw[i] = tw[i - 1] * omega % R
    chunk, tchunk = 2, n // 2
    for _ in range(log_n):
        half = chunk // 2
        for base in range(0, n, chunk):
            for i in range(half):
                t = a[base + half + i] * tw[i * tchunk] % R
                u = a[base + i]
                a[base + i] = (u + t) % R
                a[base + half + i] = (u - t) % R
        chunk *= 2
        tchunk //= 2
    return a


def dft_naive(a, omega: int):
    n = len(a)
    return [sum(a[j] * pow(omega, i * j, R) for j in range(n)) % R for i in range(n)]


class EvaluationDomain:
    """halo2_proofs::poly::EvaluationDomain::new(j, k) restated.

    j = constraint-system degree.  quotient_poly_degree = j-1; extended_k is the
    least e with 2^e >= n*(j-1).  The extended coset is zeta * <omega_ext>
    (g_coset = Fr::ZETA, g_coset_inv = zeta^2)."""

    def __init__(self, j: int, k: int):
        self.k = k
        self.n = 1 << k
        self.quotient_poly_degree = j - 1
        ek = k
        while (1 << ek) < self.n * self.quotient_poly_degree:
            ek += 1
        self.extended_k = ek
        self.extended_n = 1 << ek
        self.extended_omega = pow(FR_ROOT_OF_UNITY, 1 << (FR_S - ek), R)
        self.extended_omega_inv = pow(self.extended_omega, -1, R)
        self.omega = pow(self.extended_omega, 1 << (ek - k), R)
        self.omega_inv = pow(self.omega, -1, R)
        self.g_coset = FR_ZETA
        self.g_coset_inv = FR_ZETA * FR_ZETA % R
        self.ifft_divisor = pow(self.n, -1, R)
        self.extended_ifft_divisor = pow(self.extended_n, -1, R)
        # t_evaluations: 1/((zeta*omega_ext^i)^n - 1) for i in 0..2^(ek-k)
        m = 1 << (ek - k)
        self.t_evaluations = [
            pow((pow(self.g_coset * pow(self.extended_omega, i, R), self.n, R) - 1) % R, -1, R)
            for i in range(m)
        ]

    def lagrange_to_coeff(self, a):
        out = best_fft(a, self.omega_inv, self.k)
        return [x * self.ifft_divisor % R for x in out]

    def coeff_to_lagrange(self, a):
        return best_fft(a, self.omega, self.k)

    def coeff_to_extended(self, a):
        assert len(a) == self.n
        cp = [1, self.g_coset, self.g_coset_inv]
        b = [x * cp[i % 3] % R for i, x in enumerate(a)] + [0] * (self.extended_n - self.n)
        return best_fft(b, self.extended_omega, self.extended_k)

    def extended_to_coeff(self, a):
        assert len(a) == self.extended_n
        out = best_fft(a, self.extended_omega_inv, self.extended_k)
        cp = [1, self.g_coset_inv, self.g_coset]
        out = [x * self.extended_ifft_divisor % R * cp[i % 3] % R for i, x in enumerate(out)]
        return out[: self.n * self.quotient_poly_degree]

    def divide_by_vanishing_poly(self, a):
        m = len(self.t_evaluations)
        return [x * self.t_evaluations[i % m] % R for i, x in enumerate(a)]

    def rotate_omega(self, x: int, rot: int) -> int:
        return x * pow(self.omega, rot, R) % R


def poly_eval(coeffs, x: int) -> int:
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % R
    return acc


# --- deterministic test-vector PRNG (splitmix64), shared with the C oracle and the tests ----
class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def field(self, mod: int) -> int:
        """Uniform-ish element: 256 random bits with the top 2 cleared, reduced mod `mod`."""
        v = 0
        for i in range(4):
            v |= self.next() << (64 * i)
        v &= (1 << 254) - 1
        return v % mod


# --- evaluate_h (+ divide_by_vanishing_poly) for the ECDSA circuit shape, python ints -------------
def quotient_ecdsa(shape: dict, cols: dict, ch: dict):
    """Row-by-row restatement of halo2_proofs::plonk::evaluation::Evaluator::evaluate_h followed by
    divide_by_vanishing_poly for the FlexGate + Range constraint system of ECDSACircuit::configure
    (halo2-circuits/src/ecc/ecdsa_p256.rs:94-115).  Constraint list and folding order are those of
    the reference's generated verifier (proving-server/P256Verifier.yul:406-547).

    shape: dict(k, ext_k, num_advice, num_lookup_advice, num_fixed, blinding_factors, cs_degree)
    cols : canonical ints, lists over the extended domain (same keys as zkw_quotient_inputs)
    ch   : dict(y, beta, gamma) canonical ints."""
    k, ek = shape["k"], shape["ext_k"]
    A, L, F = shape["num_advice"], shape["num_lookup_advice"], shape["num_fixed"]
    en = 1 << ek
    rs = 1 << (ek - k)
    last = -(shape["blinding_factors"] + 1)
    chunk = shape["cs_degree"] - 2
    ncols = F + A + L
    nsets = (ncols + chunk - 1) // chunk
    nlk = L or 1
    dom = EvaluationDomain(shape["cs_degree"], k)
    assert dom.extended_k == ek
    y, beta, gamma = ch["y"], ch["beta"], ch["gamma"]
    pcols = list(cols["constants"]) + list(cols["advice"])
    out = []
    for i in range(en):
        rot = lambda r: (i + r * rs) % en
        v = 0
        for c in range(A):
            a = cols["advice"][c]
            g = cols["q_enable"][c][i] * (a[i] + a[rot(1)] * a[rot(2)] - a[rot(3)])
            v = (v * y + g) % R
        l0, ll, la = cols["l0"][i], cols["l_last"][i], cols["l_active"][i]
        z = cols["perm_z"]
        v = (v * y + (1 - z[0][i]) * l0) % R
        v = (v * y + (z[-1][i] * z[-1][i] - z[-1][i]) * ll) % R
        for s in range(1, nsets):
            v = (v * y + (z[s][i] - z[s - 1][rot(last)]) * l0) % R
        x = dom.g_coset * pow(dom.extended_omega, i, R) % R
        cur = beta * x % R
        for s in range(nsets):
            left, right = z[s][rot(1)], z[s][i]
            for c in range(s * chunk, min((s + 1) * chunk, ncols)):
                left = left * (pcols[c][i] + beta * cols["sigma"][c][i] + gamma) % R
                right = right * (pcols[c][i] + cur + gamma) % R
                cur = cur * FR_DELTA % R
            v = (v * y + (left - right) * la) % R
        for q in range(nlk):
            lz, ap, sp = cols["lookup_z"][q], cols["lookup_a"][q], cols["lookup_s"][q]
            inp = cols["advice"][A + q][i] if L else cols["q_lookup"][i] * cols["advice"][0][i] % R
            v = (v * y + (1 - lz[i]) * l0) % R
            v = (v * y + (lz[i] * lz[i] - lz[i]) * ll) % R
            lhs = lz[rot(1)] * (ap[i] + beta) % R * (sp[i] + gamma) % R
            rhs = lz[i] * (inp + beta) % R * (cols["table"][i] + gamma) % R
            v = (v * y + (lhs - rhs) * la) % R
            v = (v * y + (ap[i] - sp[i]) * l0) % R
            v = (v * y + (ap[i] - sp[i]) * (ap[i] - ap[rot(-1)]) % R * la) % R
        out.append(v * dom.t_evaluations[i % rs] % R)
    return out
