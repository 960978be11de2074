"""Deterministic P-256 signed assertions for the tests (the recipe of the reference's own circuit test,
halo2-circuits/src/ecc/ecdsa_p256.rs:222-234: random key, random message hash, r = (kG).x mod n,
s = k^-1 (m + r sk)), in the wire format the browser sends (five 32-byte little-endian encodings,
web-demo/src/pages/index.tsx:285-292).  Plain Python integers; test infrastructure only."""
import hashlib

P = 0xFFFFFFFF00000001000000000000000000000000FFFFFFFFFFFFFFFFFFFFFFFF
N = 0xFFFFFFFF00000000FFFFFFFFFFFFFFFFBCE6FAADA7179E84F3B9CAC2FC632551
A = P - 3
B = 0x5AC635D8AA3A93E7B3EBBD55769886BC651D06B0CC53B0F63BCE3C3E27D2604B
G = (0x6B17D1F2E12C4247F8BCE6E563A440F277037D812DEB33A0F4A13945D898C296,
     0x4FE342E2FE1A7F9B8EE7EB4A7C0F9E162BCE33576B315ECECBB6406837BF51F5)


def add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    if p[0] == q[0]:
        if (p[1] + q[1]) % P == 0:
            return None
        lam = (3 * p[0] * p[0] + A) * pow(2 * p[1], -1, P) % P
    else:
        lam = (q[1] - p[1]) * pow(q[0] - p[0], -1, P) % P
    x = (lam * lam - p[0] - q[0]) % P
    return x, (lam * (p[0] - x) - p[1]) % P


def mul(p, k):
    acc = None
    while k:
        if k & 1:
            acc = add(acc, p)
        p = add(p, p)
        k >>= 1
    return acc


def _h(seed, tag):
    return int.from_bytes(hashlib.sha256(b"zkw-assertion-%d-" % seed + tag).digest(), "big") % (N - 1) + 1


def signed_ints(seed: int):
    sk, k, m = _h(seed, b"sk"), _h(seed, b"k"), _h(seed, b"msg")
    pk = mul(G, sk)
    r = mul(G, k)[0] % N
    s = pow(k, -1, N) * (m + r * sk) % N
    assert r and s
    return {"pubkey_x": pk[0], "pubkey_y": pk[1], "r": r, "s": s, "msg_hash": m}


def signed_assertion(seed: int) -> dict:
    return {k: v.to_bytes(32, "little") for k, v in signed_ints(seed).items()}


def verify_ints(a) -> bool:
    r, s, m = a["r"], a["s"], a["msg_hash"]
    if not (0 < r < N and 0 < s < N):
        return False
    w = pow(s, -1, N)
    pt = add(mul(G, m * w % N), mul((a["pubkey_x"], a["pubkey_y"]), r * w % N))
    return pt is not None and pt[0] % N == r
