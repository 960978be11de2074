"""Host-side circuit description for the P-256 ECDSA prover: the column layout that
ECDSACircuit::configure produces (halo2-circuits/src/ecc/ecdsa_p256.rs:94-115, from the JSON configs under
halo2-circuits/src/configs/) and a satisfying assignment with that layout.

halo2-ecc's actual ECDSA cell layout lives in an un-vendored crate (halo2-circuits/Cargo.toml:12-13) and
cannot be reproduced here, so the assignment below is a *shape-identical synthetic witness*: the same
vertical gate q*(a + b*c - d), the same range-lookup and the same kinds of copy constraints (gate chaining,
constants, lookup cells), filled from a PRNG keyed by the assertion (r, s, msghash, public key).  The
prover work per proof — commitments, NTTs, quotient, openings — is the same as for the real witness; the
statement proven is satisfaction of this synthetic system, not ECDSA validity.  See DESIGN.md.
"""
from __future__ import annotations

import hashlib
import json
from dataclasses import dataclass

import numpy as np

# secp256r1 parameters (input validation mirrors Fp::from_bytes / Fq::from_bytes / from_xy at ecdsa_p256.rs:346-352)
P256_P = 0xFFFFFFFF00000001000000000000000000000000FFFFFFFFFFFFFFFFFFFFFFFF
P256_N = 0xFFFFFFFF00000000FFFFFFFFFFFFFFFFBCE6FAADA7179E84F3B9CAC2FC632551
P256_B = 0x5AC635D8AA3A93E7B3EBBD55769886BC651D06B0CC53B0F63BCE3C3E27D2604B


@dataclass
class CircuitParams:
    """The one-line JSON config of the reference (struct CircuitParams, ecdsa_p256.rs:53-63)."""
    strategy: str = "Simple"
    degree: int = 17
    num_advice: int = 4
    num_lookup_advice: int = 1
    num_fixed: int = 1
    lookup_bits: int = 16
    limb_bits: int = 88
    num_limbs: int = 3

    @classmethod
    def from_json(cls, text: str) -> "CircuitParams":
        return cls(**json.loads(text))

    @classmethod
    def for_degree(cls, degree: int) -> "CircuitParams":
        """The nine lines of halo2-circuits/src/configs/bench_ecdsa.config."""
        table = {19: (1, 1, 1, 18, 88), 18: (2, 1, 1, 17, 88), 17: (4, 1, 1, 16, 88), 16: (8, 2, 1, 15, 90), 15: (17, 3, 1, 14, 90),
                 14: (34, 6, 1, 13, 91), 13: (68, 12, 1, 12, 88), 12: (139, 24, 2, 11, 88), 11: (291, 53, 4, 10, 88)}
        if degree not in table:
            raise ValueError(f"no reference config for degree {degree}")
        a, l, f, lb, limb = table[degree]
        return cls("Simple", degree, a, l, f, lb, limb, 3)


def validate_assertion(pubkey_x: bytes, pubkey_y: bytes, r: bytes, s: bytes, msg_hash: bytes) -> tuple[int, int, int, int, int]:
    """Little-endian canonical encodings, as the browser sends them (web-demo/src/pages/index.tsx:285-292).
    Raises ValueError where the reference's `.unwrap()` on from_bytes / from_xy would panic."""
    vals = []
    for name, b, mod in (("pubkey_x", pubkey_x, P256_P), ("pubkey_y", pubkey_y, P256_P), ("r", r, P256_N), ("s", s, P256_N),
                         ("msg_hash", msg_hash, P256_N)):
        if len(b) != 32:
            raise ValueError(f"{name}: expected 32 bytes")
        v = int.from_bytes(bytes(b), "little")
        if v >= mod:
            raise ValueError(f"{name}: non-canonical field element")
        vals.append(v)
    x, y = vals[0], vals[1]
    if (y * y - (x * x * x - 3 * x + P256_B)) % P256_P != 0:
        raise ValueError("public key is not on secp256r1")
    return tuple(vals)


class InvalidSignature(ValueError):
    """The assertion's (r, s) is not a valid P-256 ECDSA signature of msg_hash under the public key: the circuit has
    no satisfying assignment, so no proof is attempted."""


class EcdsaCircuit:
    """The P-256 ECDSA verification circuit (ECDSACircuit, ecdsa_p256.rs:65-207) for one config line: layout and
    witness synthesis live in the library (csrc/ecdsa_circuit.cpp); this class is the host-side handle.

    fixed_columns() / permutation_mapping() feed keygen (the `without_witnesses` pass, ecdsa_p256.rs:90-92,256-260);
    synthesize() is ECDSACircuit::synthesize (:117-206) for one assertion."""

    def __init__(self, params: CircuitParams):
        import ctypes as C
        from . import native
        self.params = params
        self._lib = native.load_library()
        cp = native.CircuitParamsC(params.degree, params.num_advice, params.num_lookup_advice, params.num_fixed, params.lookup_bits,
                                   params.limb_bits, params.num_limbs)
        h = C.c_void_p()
        rc = self._lib.zkw_ecdsa_circuit_new(C.byref(cp), C.byref(h))
        self._h = h
        self.shape = native.CircuitShape()
        if h:
            self._lib.zkw_ecdsa_circuit_shape(h, C.byref(self.shape))
        if rc == native.ZKW_ERR_UNSUPPORTED:
            raise ValueError(f"the ECDSA circuit does not fit config {params} ({self.stats()})")
        if rc != native.ZKW_OK:
            raise native.ZkwError(rc, "zkw_ecdsa_circuit_new")
        self.k, self.n = params.degree, 1 << params.degree
        self.A, self.F, self.L = self.shape.num_advice, self.shape.num_fixed, self.shape.num_lookup_advice
        self.selector_mode = self.L == 0
        self.nfixed = self.F + 1 + self.A + (1 if self.selector_mode else 0)
        self.nperm = self.F + self.A + self.L
        self.rows = self._rows()[0]

    def _rows(self):
        import ctypes as C
        ncol = self.shape.num_advice + self.shape.num_lookup_advice
        rows = (C.c_size_t * ncol)()
        stats = (C.c_uint64 * 4)()
        self._lib.zkw_ecdsa_circuit_rows(self._h, rows, stats)
        return [int(x) for x in rows], [int(x) for x in stats]

    def stats(self) -> dict:
        rows, st = self._rows()
        return {"rows": rows, "gate_cells": st[0], "lookups": st[1], "constants": st[2], "copies": st[3]}

    def fixed_columns(self) -> list[np.ndarray]:
        """canonical integers, (n, 4) uint64 each: [constants.., table, q_enable.., (q_lookup)]"""
        import ctypes as C
        from . import native
        cols = [np.zeros((self.n, 4), dtype=np.uint64) for _ in range(self.nfixed)]
        ptrs = (native.u64p * self.nfixed)(*[c.ctypes.data_as(native.u64p) for c in cols])
        rc = self._lib.zkw_ecdsa_circuit_fixed(self._h, ptrs)
        if rc != native.ZKW_OK:
            raise native.ZkwError(rc, "zkw_ecdsa_circuit_fixed")
        return cols

    def permutation_mapping(self) -> list[np.ndarray]:
        import ctypes as C
        from . import native
        u32p = C.POINTER(C.c_uint32)
        maps = [np.zeros((self.n, 2), dtype=np.uint32) for _ in range(self.nperm)]
        ptrs = (u32p * self.nperm)(*[m.ctypes.data_as(u32p) for m in maps])
        rc = self._lib.zkw_ecdsa_circuit_permutation(self._h, ptrs)
        if rc != native.ZKW_OK:
            raise native.ZkwError(rc, "zkw_ecdsa_circuit_permutation")
        return maps

    def synthesize(self, pubkey_x: bytes, pubkey_y: bytes, r: bytes, s: bytes, msg_hash: bytes, out: list[np.ndarray] | None = None,
                   allow_invalid: bool = False) -> list[np.ndarray]:
        """Advice columns (gate columns, then lookup-advice columns) as (rows, 4) uint64 canonical integers, written
        into `out` (e.g. views of page-locked memory) or fresh arrays.  Raises InvalidSignature unless the signature
        verifies (the reference would emit a proof of an unsatisfied system instead, ecdsa_p256.rs:182-191)."""
        import ctypes as C
        from . import native
        if out is None:
            out = [np.zeros((r_, 4), dtype=np.uint64) for r_ in self.rows]
        if len(out) != len(self.rows) or any(o.dtype != np.uint64 or o.shape != (r_, 4) or not o.flags.c_contiguous for o, r_ in zip(out, self.rows)):
            raise ValueError("synthesize: out must be C-contiguous (rows, 4) uint64 arrays, rows = EcdsaCircuit.rows")
        ptrs = (native.u64p * len(out))(*[o.ctypes.data_as(native.u64p) for o in out])
        ok = C.c_int(0)
        rc = self._lib.zkw_ecdsa_synthesize(self._h, bytes(pubkey_x), bytes(pubkey_y), bytes(r), bytes(s), bytes(msg_hash), ptrs, None, C.byref(ok))
        if rc != native.ZKW_OK:
            raise native.ZkwError(rc, "zkw_ecdsa_synthesize")
        if not ok.value and not allow_invalid:
            raise InvalidSignature("signature does not verify: the ECDSA circuit has no satisfying assignment")
        return out

    def close(self):
        if getattr(self, "_h", None):
            self._lib.zkw_ecdsa_circuit_free(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


def _u64_to_limbs(v: np.ndarray) -> np.ndarray:
    out = np.zeros((v.shape[0], 4), dtype=np.uint64)
    out[:, 0] = v
    return out


class SyntheticEcdsaCircuit:
    """Layout: gate g of gate column c occupies rows 4g..4g+3 (a, b, c, d = a + b*c), q_enable_c[4g] = 1.
    Odd gates chain a_g = d_{g-1} (copy constraint); the first eight even gates of column 0 take their `a`
    from the constants column (copy constraint); b < 2^lookup_bits is range-checked: in selector mode
    q_lookup[4g+1] = 1, otherwise every third b of column 0 is copied into lookup-advice column 0 and the
    rest of the lookup columns hold in-range values.  Values are canonical integers < 2^64."""

    def __init__(self, params: CircuitParams, blinding_factors: int = 6):
        self.params = params
        self.k = params.degree
        self.n = 1 << self.k
        self.u = self.n - (blinding_factors + 1)
        self.A, self.F = params.num_advice, params.num_fixed
        self.selector_mode = params.num_advice == 1
        self.L = 0 if self.selector_mode else params.num_lookup_advice
        self.T = min(1 << params.lookup_bits, self.u)
        self.G = self.u // 4                      # gates per gate column: rows 4g..4g+3 < u
        self.nfixed = self.F + 1 + self.A + (1 if self.selector_mode else 0)
        self.nperm = self.F + self.A + self.L
        self.table_col = self.F
        self.q_lookup_col = self.F + 1 + self.A

    # -- keygen side -------------------------------------------------------------------------------
    def fixed_columns(self) -> list[np.ndarray]:
        """canonical values, (n,) uint64 each: [constants.., table, q_enable.., (q_lookup)]"""
        n, G = self.n, self.G
        cols = [np.zeros(n, dtype=np.uint64) for _ in range(self.nfixed)]
        for f in range(self.F):
            j = np.arange(min(8, self.u), dtype=np.uint64)
            cols[f][: j.shape[0]] = (j + 1) * np.uint64(f + 1)
        cols[self.table_col][: self.T] = np.arange(self.T, dtype=np.uint64)
        rows = 4 * np.arange(G)
        for c in range(self.A):
            cols[self.F + 1 + c][rows] = 1
        if self.selector_mode:
            cols[self.q_lookup_col][rows + 1] = 1
        return cols

    def copy_pairs(self) -> list[tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]]:
        """(col_a, row_a, col_b, row_b) arrays of equality constraints between permutation columns
        [constants.., gate advice.., lookup advice..]; every cell appears in at most one pair."""
        G, F = self.G, self.F
        out = []
        g_odd = np.arange(1, G, 2)
        for c in range(self.A):
            col = np.full(g_odd.shape[0], F + c, dtype=np.uint32)
            out.append((col, (4 * g_odd).astype(np.uint32), col, (4 * g_odd - 1).astype(np.uint32)))
        nconst = min(8, (G + 1) // 2, self.u)
        j = np.arange(nconst)
        out.append((np.zeros(nconst, dtype=np.uint32), j.astype(np.uint32), np.full(nconst, F, dtype=np.uint32), (8 * j).astype(np.uint32)))
        if self.L:
            jj = np.arange(0, min(G, self.u), 3)
            out.append((np.full(jj.shape[0], F, dtype=np.uint32), (4 * jj + 1).astype(np.uint32),
                        np.full(jj.shape[0], F + self.A, dtype=np.uint32), jj.astype(np.uint32)))
        return out

    def permutation_mapping(self) -> list[np.ndarray]:
        """Per permutation column an (n, 2) uint32 array of (col', row'): halo2's Assembly after the copy
        constraints above (merging two singleton cycles swaps their successors)."""
        n = self.n
        maps = []
        for c in range(self.nperm):
            m = np.empty((n, 2), dtype=np.uint32)
            m[:, 0] = c
            m[:, 1] = np.arange(n, dtype=np.uint32)
            maps.append(m)
        for ca, ra, cb, rb in self.copy_pairs():
            for i in range(ca.shape[0]):
                a, b = (int(ca[i]), int(ra[i])), (int(cb[i]), int(rb[i]))
                maps[a[0]][a[1]] = b
                maps[b[0]][b[1]] = a
        return maps

    # -- witness side --------------------------------------------------------------------------------
    def synthesize(self, assertion: bytes) -> list[np.ndarray]:
        """Advice columns' usable rows as canonical (rows,) uint64 arrays, keyed by the assertion bytes.
        numpy statement of the generator that the library's host routine zkw_synth_witness (csrc/witness.cu)
        runs inside the timed end-to-end path; tests/test_circuit_cpu.py checks the two agree cell for cell."""
        G, u, T = self.G, self.u, self.T
        seed = int.from_bytes(hashlib.blake2b(b"zkw-b200-synth" + assertion).digest()[:8], "little")
        pow2 = T & (T - 1) == 0
        U = np.uint64

        def mix64(z):   # splitmix64 finaliser over a counter (uint64 arithmetic wraps)
            z = (z ^ (z >> U(30))) * U(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> U(27))) * U(0x94D049BB133111EB)
            return z ^ (z >> U(31))

        def stream(col_index):
            return U((seed ^ (0x9E3779B97F4A7C15 * (col_index + 1))) & 0xFFFFFFFFFFFFFFFF)

        cols = []
        with np.errstate(over="ignore"):
            g3 = U(3) * np.arange(G, dtype=np.uint64)
            for c in range(self.A):
                s = stream(c)
                r1, r2, r3 = mix64(s + g3), mix64(s + g3 + U(1)), mix64(s + g3 + U(2))
                # a: a mix of bits and wide (62-bit) limbs; b < T is the range-checked cell; c < 2^40
                a = np.where(r2 >> U(63) != 0, (r2 >> U(62)) & U(1), r1 >> U(2))
                b = (r3 & U(T - 1)) if pow2 else (r3 % U(T))
                cc = (r2 >> U(8)) & U((1 << 40) - 1)
                if c == 0:
                    nconst = min(8, (G + 1) // 2)
                    a[0:2 * nconst:2] = np.arange(1, nconst + 1, dtype=np.uint64)
                bc = b * cc
                # even gates are free, odd gates take a = d of the previous (even) gate
                d_even = a[0::2] + bc[0::2]
                if G > 1:
                    a[1::2] = d_even[: a[1::2].shape[0]]
                col = np.empty(4 * G, dtype=np.uint64)
                quad = col.reshape(G, 4)
                quad[:, 0], quad[:, 1], quad[:, 2], quad[:, 3] = a, b, cc, a + bc
                cols.append(col)
            for l in range(self.L):
                r = mix64(stream(self.A + l) + np.arange(u, dtype=np.uint64))
                col = (r & U(T - 1)) if pow2 else (r % U(T))
                if l == 0:
                    jj = np.arange(0, min(G, u), 3)
                    col[jj] = cols[0][4 * jj + 1]
                cols.append(col)
        return cols


def to_limbs(col: np.ndarray) -> np.ndarray:
    return _u64_to_limbs(np.ascontiguousarray(col, dtype=np.uint64))
