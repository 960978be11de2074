"""BN254 optimal-ate pairing in plain Python integers.  TEST INFRASTRUCTURE ONLY.

Needed to run the reference's generated verifier (proving-server/P256Verifier.yul) on the
reference's golden proof (contracts/test/P256Account.t.sol:120): the Yul ends in a call to the EVM
pairing precompile (0x08, yul:1125-1136).  Fp12 is represented as Fp[w]/(w^12 - 18 w^6 + 82)
(i = w^6 - 9 satisfies i^2 = -1, xi = 9 + i = w^6), the usual flat representation: slow, simple.
"""
from __future__ import annotations

from .pyref import P, R

ATE_LOOP_COUNT = 29793968203157093288
LOG_ATE_LOOP_COUNT = 63
FQ12_MOD = [82, 0, 0, 0, 0, 0, -18, 0, 0, 0, 0, 0]  # w^12 = 18 w^6 - 82


class FQ12:
    __slots__ = ("c",)

    def __init__(self, c):
        self.c = [x % P for x in c]
        assert len(self.c) == 12

    @classmethod
    def one(cls):
        return cls([1] + [0] * 11)

    @classmethod
    def zero(cls):
        return cls([0] * 12)

    @classmethod
    def from_int(cls, x):
        return cls([x] + [0] * 11)

    def __add__(self, o):
        return FQ12([a + b for a, b in zip(self.c, o.c)])

    def __sub__(self, o):
        return FQ12([a - b for a, b in zip(self.c, o.c)])

    def __neg__(self):
        return FQ12([-a for a in self.c])

    def __eq__(self, o):
        return self.c == o.c

    def scale(self, k: int):
        return FQ12([a * k for a in self.c])

    def __mul__(self, o):
        if isinstance(o, int):
            return self.scale(o)
        t = [0] * 23
        for i, a in enumerate(self.c):
            if a == 0:
                continue
            for j, b in enumerate(o.c):
                t[i + j] += a * b
        # reduce: w^12 = 18 w^6 - 82
        for k in range(22, 11, -1):
            v = t[k]
            if v:
                t[k - 6] += 18 * v
                t[k - 12] -= 82 * v
        return FQ12(t[:12])

    def __pow__(self, e: int):
        result = FQ12.one()
        base = self
        while e:
            if e & 1:
                result = result * base
            base = base * base
            e >>= 1
        return result

    def is_one(self):
        return self.c == [1] + [0] * 11


# ---- G2 over Fp2, embedded into Fp12 through the twist ---------------------------------------------
def fq2_to_fq12(c0: int, c1: int) -> FQ12:
    """a = c0 + c1 * i with i = w^6 - 9."""
    return FQ12([c0 - 9 * c1, 0, 0, 0, 0, 0, c1, 0, 0, 0, 0, 0])


W = FQ12([0, 1] + [0] * 10)
W3 = W * W * W


def fq2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def fq2_inv(a):
    d = pow(a[0] * a[0] + a[1] * a[1], -1, P)
    return (a[0] * d % P, -a[1] * d % P)


def fq2_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def fq2_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def fq2_pow(a, e: int):
    r = (1, 0)
    while e:
        if e & 1:
            r = fq2_mul(r, a)
        a = fq2_mul(a, a)
        e >>= 1
    return r


XI = (9, 1)
FROB_X = fq2_pow(XI, (P - 1) // 3)   # w^(2p) = w^2 * xi^((p-1)/3)
FROB_Y = fq2_pow(XI, (P - 1) // 2)   # w^(3p) = w^3 * xi^((p-1)/2)


def _frobenius_twist(q):
    (x0, x1), (y0, y1) = q
    return (fq2_mul((x0, -x1 % P), FROB_X), fq2_mul((y0, -y1 % P), FROB_Y))


def _line(t, q, p):
    """Line through twist points t, q (tangent if equal) evaluated at the G1 point p, as an element
    of Fp12: with the untwist (x', y') -> (x' w^2, y' w^3) the chord of slope lam' becomes
    l(P) = yP - lam' xP w + (lam' x_t' - y_t') w^3.  Returns (value, t + q)."""
    (x1, y1), (x2, y2) = t, q
    if x1 == x2 and y1 == y2:
        lam = fq2_mul(fq2_mul((3, 0), fq2_mul(x1, x1)), fq2_inv(fq2_mul((2, 0), y1)))
    else:
        assert x1 != x2, "vertical line in the Miller loop"
        lam = fq2_mul(fq2_sub(y2, y1), fq2_inv(fq2_sub(x2, x1)))
    x3 = fq2_sub(fq2_sub(fq2_mul(lam, lam), x1), x2)
    y3 = fq2_sub(fq2_mul(lam, fq2_sub(x1, x3)), y1)
    xp, yp = p
    c1 = fq2_mul(lam, (-xp % P, 0))
    c3 = fq2_sub(fq2_mul(lam, x1), y1)
    val = FQ12.from_int(yp) + fq2_to_fq12(*c1) * W + fq2_to_fq12(*c3) * W3
    return val, (x3, y3)


def miller_loop(q, p) -> FQ12:
    """q: G2 affine over Fp2 (twist coordinates), p: G1 affine; either None -> 1."""
    if q is None or p is None:
        return FQ12.one()
    r = q
    f = FQ12.one()
    for i in range(LOG_ATE_LOOP_COUNT, -1, -1):
        l, r = _line(r, r, p)
        f = f * f * l
        if ATE_LOOP_COUNT & (1 << i):
            l, r = _line(r, q, p)
            f = f * l
    q1 = _frobenius_twist(q)
    q2 = _frobenius_twist(q1)
    nq2 = (q2[0], (-q2[1][0] % P, -q2[1][1] % P))
    l, r = _line(r, q1, p)
    f = f * l
    l, r = _line(r, nq2, p)
    f = f * l
    return f


def final_exponentiate(f: FQ12) -> FQ12:
    return f ** ((P ** 12 - 1) // R)


def pairing_product_is_one(pairs) -> bool:
    """pairs: iterable of (G1 affine (x,y) | None, G2 affine ((x0,x1),(y0,y1)) | None): the EVM 0x08
    precompile's predicate  prod e(P_i, Q_i) == 1."""
    f = FQ12.one()
    for p, q in pairs:
        f = f * miller_loop(q, p)
    return final_exponentiate(f).is_one()


# ---- G2 arithmetic over Fp2 (for subgroup / known-answer checks) ------------------------------------
G2_B = fq2_mul((3, 0), fq2_inv((9, 1)))
G2_GEN = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)


def g2_is_on_curve(q) -> bool:
    if q is None:
        return True
    x, y = q
    return fq2_sub(fq2_mul(y, y), fq2_add(fq2_mul(fq2_mul(x, x), x), G2_B)) == (0, 0)


def g2_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    (x1, y1), (x2, y2) = a, b
    if x1 == x2:
        if fq2_add(y1, y2) == (0, 0):
            return None
        lam = fq2_mul(fq2_mul((3, 0), fq2_mul(x1, x1)), fq2_inv(fq2_mul((2, 0), y1)))
    else:
        lam = fq2_mul(fq2_sub(y2, y1), fq2_inv(fq2_sub(x2, x1)))
    x3 = fq2_sub(fq2_sub(fq2_mul(lam, lam), x1), x2)
    y3 = fq2_sub(fq2_mul(lam, fq2_sub(x1, x3)), y1)
    return (x3, y3)


def g2_mul(a, k: int):
    acc = None
    while k:
        if k & 1:
            acc = g2_add(acc, a)
        a = g2_add(a, a)
        k >>= 1
    return acc
