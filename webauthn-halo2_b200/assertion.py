"""Synthetic WebAuthn assertions for benchmarks and demos: a deterministic P-256 key and a
signature over  msg_hash = SHA-256(authenticatorData || SHA-256(clientDataJSON))  — the value the browser computes
(web-demo/src/pages/index.tsx:186-197) — produced the way the reference's own circuit test signs
(halo2-circuits/src/ecc/ecdsa_p256.rs:222-234: r = (kG).x mod n, s = k^-1 (m + r sk)), encoded as the five 32-byte
little-endian values generate_proof{,_evm} take (ecdsa_p256.rs:329,379; index.tsx:285-292).

Plain Python integers (a few milliseconds per assertion): this is input generation, not part of any timed path."""
from __future__ import annotations

import hashlib

from .circuit import P256_B, P256_N, P256_P

P256_G = (0x6B17D1F2E12C4247F8BCE6E563A440F277037D812DEB33A0F4A13945D898C296,
          0x4FE342E2FE1A7F9B8EE7EB4A7C0F9E162BCE33576B315ECECBB6406837BF51F5)


def _add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    if p[0] == q[0]:
        if (p[1] + q[1]) % P256_P == 0:
            return None
        lam = (3 * p[0] * p[0] - 3) * pow(2 * p[1], -1, P256_P) % P256_P
    else:
        lam = (q[1] - p[1]) * pow(q[0] - p[0], -1, P256_P) % P256_P
    x = (lam * lam - p[0] - q[0]) % P256_P
    return x, (lam * (p[0] - x) - p[1]) % P256_P


def _mul(p, k):
    acc = None
    while k:
        if k & 1:
            acc = _add(acc, p)
        p = _add(p, p)
        k >>= 1
    return acc


def _h(seed: int, tag: bytes) -> int:
    return int.from_bytes(hashlib.sha256(b"zkw-b200-assertion|%d|" % seed + tag).digest(), "big") % (P256_N - 1) + 1


def webauthn_message_hash(seed: int) -> int:
    """SHA-256(authData || SHA-256(clientDataJSON)) for a 37-byte authenticatorData and the fixed JSON template."""
    rp_id_hash = hashlib.sha256(b"zkwebauthn.example").digest()
    auth_data = rp_id_hash + b"\x05" + (seed & 0xFFFFFFFF).to_bytes(4, "big")           # flags UP|UV, signature counter
    challenge = hashlib.sha256(b"challenge|%d" % seed).hexdigest()
    client_data = ('{"type":"webauthn.get","challenge":"%s","origin":"https://zkwebauthn.example","crossOrigin":false}' % challenge).encode()
    return int.from_bytes(hashlib.sha256(auth_data + hashlib.sha256(client_data).digest()).digest(), "big") % P256_N


def synthetic_assertion(seed: int) -> bytes:
    """160 bytes: pubkey_x | pubkey_y | r | s | msg_hash, each little-endian."""
    sk, k = _h(seed, b"sk"), _h(seed, b"k")
    m = webauthn_message_hash(seed)
    pk = _mul(P256_G, sk)
    r = _mul(P256_G, k)[0] % P256_N
    s = pow(k, -1, P256_N) * (m + r * sk) % P256_N
    if r == 0 or s == 0:      # pragma: no cover - probability 2^-256
        return synthetic_assertion(seed + (1 << 32))
    assert (pk[1] * pk[1] - (pk[0] ** 3 - 3 * pk[0] + P256_B)) % P256_P == 0
    return b"".join(v.to_bytes(32, "little") for v in (pk[0], pk[1], r, s, m))
