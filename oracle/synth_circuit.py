"""Synthetic ECDSA-shaped circuit instances for the oracle-side prover tests: a satisfying assignment
for the FlexGate + Range constraint system (same columns, gates, lookup and permutation structure as
ECDSACircuit::configure, halo2-circuits/src/ecc/ecdsa_p256.rs:94-115) without halo2-ecc's actual
ECDSA cell layout, which lives in an un-vendored crate.  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import random

from .halo2_ref import Shape
from .pyref import R


class Assembly:
    """halo2_proofs::plonk::permutation::keygen::Assembly restated: cycles as a successor mapping."""

    def __init__(self, ncols, n):
        self.n = n
        self.mapping = [[(c, r) for r in range(n)] for c in range(ncols)]
        self.aux = [[(c, r) for r in range(n)] for c in range(ncols)]
        self.sizes = [[1] * n for _ in range(ncols)]

    def copy(self, left, right):
        (lc, lr), (rc, rr) = left, right
        if self.aux[lc][lr] == self.aux[rc][rr]:
            return
        la, ra = self.aux[lc][lr], self.aux[rc][rr]
        if self.sizes[la[0]][la[1]] < self.sizes[ra[0]][ra[1]]:
            left, right = right, left
            (lc, lr), (rc, rr) = left, right
            la, ra = ra, la
        self.sizes[la[0]][la[1]] += self.sizes[ra[0]][ra[1]]
        i = ra
        while True:
            self.aux[i[0]][i[1]] = la
            i = self.mapping[i[0]][i[1]]
            if i == ra:
                break
        self.mapping[lc][lr], self.mapping[rc][rr] = self.mapping[rc][rr], self.mapping[lc][lr]


def build(shape: Shape, seed: int = 0, lookup_bits: int | None = None):
    """Returns (fixed_values, mapping, advice_usable_rows)."""
    rng = random.Random(seed)
    n, u = shape.n, shape.usable_rows
    A, L, F = shape.num_advice, shape.num_lookup_advice, shape.num_fixed
    if lookup_bits is None:
        lookup_bits = max(1, shape.k - 1)
    T = min(1 << lookup_bits, u)
    fixed = [[0] * n for _ in range(shape.num_fixed_cols)]
    for i in range(T):
        fixed[shape.table_col][i] = i
    for f in range(F):
        for j in range(min(8, u)):
            fixed[f][j] = (j + 1) * (f + 1)
    advice = [[0] * u for _ in range(A + L)]
    asm = Assembly(F + A + L, n)
    pcol = lambda kind, c: c if kind == "fixed" else F + c
    for c in range(A):
        prev_out = None
        g = 0
        for i in range(0, u - 3, 4):
            a = rng.randrange(1 << 88) if rng.random() < 0.5 else rng.randrange(2)
            if g % 2 == 1 and prev_out is not None:
                a = prev_out
                asm.copy((pcol("advice", c), i), (pcol("advice", c), i - 1))
            elif g % 2 == 0 and g // 2 < min(8, u):
                a = fixed[c % F][g // 2]
                asm.copy((pcol("fixed", c % F), g // 2), (pcol("advice", c), i))
            b = rng.randrange(T)
            cc = rng.randrange(1 << 64)
            d = (a + b * cc) % R
            advice[c][i:i + 4] = [a, b, cc, d]
            fixed[shape.q_enable_col(c)][i] = 1
            prev_out = d
            g += 1
    if shape.selector_mode:
        for i in range(1, u - 3, 4):
            fixed[shape.q_lookup_col][i] = 1     # a_0[i] = b < T by construction
    else:
        for l in range(L):
            col = A + l
            for j in range(u):
                src = 4 * j + 1
                if src < u - 3 and src // 4 * 4 + 3 < u and l == 0 and j % 3 == 0:
                    advice[col][j] = advice[0][src]
                    asm.copy((pcol("advice", 0), src), (pcol("advice", col), j))
                else:
                    advice[col][j] = rng.randrange(T)
    return fixed, asm.mapping, advice


def check_satisfied(shape: Shape, fixed, mapping, advice_usable):
    """MockProver-style check of gates, lookups and copy constraints on the usable rows."""
    u = shape.usable_rows
    A = shape.num_advice
    adv = [list(c) + [0] * (u - len(c)) for c in advice_usable]
    for c in range(A):
        for i in range(u):
            if fixed[shape.q_enable_col(c)][i]:
                assert i + 3 < u
                assert (adv[c][i] + adv[c][i + 1] * adv[c][i + 2] - adv[c][i + 3]) % R == 0, ("gate", c, i)
    table = set(fixed[shape.table_col][:u])
    for l in range(shape.num_lookups):
        for i in range(u):
            v = fixed[shape.q_lookup_col][i] * adv[0][i] % R if shape.selector_mode else adv[A + l][i]
            assert v in table, ("lookup", l, i)
    F = shape.num_fixed
    val = lambda c, r: fixed[c][r] if c < F else (adv[c - F][r] if r < u else None)
    for c, col in enumerate(mapping):
        for r, (cc, rr) in enumerate(col):
            if (cc, rr) != (c, r):
                assert r < u and rr < u and val(c, r) == val(cc, rr), ("copy", c, r, cc, rr)
    return True
