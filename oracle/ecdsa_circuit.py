"""TEST INFRASTRUCTURE ONLY — CPU restatement (plain Python integers) of the P-256 ECDSA verification circuit
that the product synthesises in C++ (webauthn-halo2_b200/csrc/ecdsa_circuit.cpp).

What it restates.  The reference's circuit (halo2-circuits/src/ecc/ecdsa_p256.rs:117-206) witnesses r, s,
msghash and the public key and calls halo2-ecc's `ecdsa_verify_no_pubkey_check` with window widths 4, 4
(:182-191) over halo2-lib's FlexGate / Range / CRT-bigint chips (limb_bits 88, num_limbs 3, lookup_bits from
the JSON config, halo2-circuits/src/configs/*.config).  halo2-lib / halo2-ecc are un-vendored git dependencies
(halo2-circuits/Cargo.toml:12-13), so their exact cell layout cannot be reproduced; what the reference DOES fix
is the constraint system every cell must live in (proving-server/P256Verifier.yul:406-547):

    gate       q_c * (a_c(X) + a_c(wX) * a_c(w^2 X) - a_c(w^3 X))       one per gate advice column
    lookup     lookup advice column (or q_lookup * a_0 when there is one gate column)  in  [0, 2^lookup_bits)
    copies     permutation over [constants.., gate advice.., lookup advice..]

This file lays out, in that system, the same computation with the same chip parameters:

    u1 = msghash / s, u2 = r / s  (mod n);  R = u1*G + u2*PK;  R.x == r;  1 <= r, s < n;  u1, u2 < n

CRT big integers (3 limbs of limb_bits, native value mod the BN254 scalar field), range checks by
decomposition into lookup_bits-wide limbs, fixed-base multiplication by 4-bit windows over constant tables,
variable-base multiplication by 4-bit windows over a 16-entry table of multiples of PK.

Deviations from halo2-ecc, stated:
  * own cell layout (every relation `sum of products = q*p` is checked with all-positive left/right chains and
    offset carries, instead of halo2-ecc's signed no-carry intermediates);
  * the point accumulators carry constant offset points with unknown discrete logarithm (hash-derived), so no
    identity / `is_started` selection logic is needed; the offsets cancel in the final addition;
  * the verification result is CONSTRAINED (R.x limbs are copy-constrained to r's), whereas the reference
    computes a result bit and never asserts it (ecdsa_p256.rs:182-191 drops the return value): an invalid
    signature has no satisfying assignment here;
  * the public key is checked to be on the curve.

`check()` is the MockProver analogue (ecdsa_p256.rs:245-247): every gate row, every lookup cell and every copy
constraint of the assignment."""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617   # BN254 scalar field

P256_P = 0xFFFFFFFF00000001000000000000000000000000FFFFFFFFFFFFFFFFFFFFFFFF
P256_N = 0xFFFFFFFF00000000FFFFFFFFFFFFFFFFBCE6FAADA7179E84F3B9CAC2FC632551
P256_B = 0x5AC635D8AA3A93E7B3EBBD55769886BC651D06B0CC53B0F63BCE3C3E27D2604B
P256_G = (0x6B17D1F2E12C4247F8BCE6E563A440F277037D812DEB33A0F4A13945D898C296,
          0x4FE342E2FE1A7F9B8EE7EB4A7C0F9E162BCE33576B315ECECBB6406837BF51F5)
WINDOW = 4          # ecdsa_p256.rs:189-190: fixed_window_bits = var_window_bits = 4
Q_OFFSET_BITS = 258  # quotients live in (-2^258, 2^258); witnessed as q + 2^258


# ---- plain P-256 arithmetic (affine, None = identity) ---------------------------------------------------
def ec_add(p, q):
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


def ec_neg(p):
    return None if p is None else (p[0], (-p[1]) % P256_P)


def ec_mul(p, k):
    acc = None
    k %= P256_N
    while k:
        if k & 1:
            acc = ec_add(acc, p)
        p = ec_add(p, p)
        k >>= 1
    return acc


def offset_point(tag: bytes):
    """A curve point nobody knows the discrete logarithm of: x = SHA-256(tag || counter) for the first counter
    that lands on the curve, the root with even y (p = 3 mod 4: sqrt = a^((p+1)/4))."""
    ctr = 0
    while True:
        x = int.from_bytes(hashlib.sha256(tag + ctr.to_bytes(4, "big")).digest(), "big") % P256_P
        rhs = (x * x * x - 3 * x + P256_B) % P256_P
        y = pow(rhs, (P256_P + 1) // 4, P256_P)
        if y * y % P256_P == rhs:
            return (x, y if y % 2 == 0 else P256_P - y)
        ctr += 1


OFFSET_VAR = offset_point(b"zkw-b200 ecdsa variable-base offset")     # B2: added to every table entry d*PK
OFFSET_FIX = offset_point(b"zkw-b200 ecdsa fixed-base offset")        # B3: 2^w * B3 added to window w's table


@dataclass
class Params:
    k: int
    num_advice: int
    num_lookup_advice: int      # as in the JSON config; ignored (selector mode) when num_advice == 1
    num_fixed: int
    lookup_bits: int
    limb_bits: int = 88
    num_limbs: int = 3
    blinding_factors: int = 6

    @property
    def selector_mode(self):
        return self.num_advice == 1

    @property
    def L(self):
        return 0 if self.selector_mode else self.num_lookup_advice

    @property
    def n(self):
        return 1 << self.k

    @property
    def usable(self):
        return self.n - (self.blinding_factors + 1)


@dataclass
class Elem:
    """CRT integer: three limb cells, the native cell, and the values."""
    limbs: list
    native: int
    value: int
    lv: list = field(default_factory=list)


class DoesNotFit(Exception):
    pass


class Builder:
    """Cell allocator + chips.  A cell id is col * n + row over the advice columns [gate.., lookup..]."""

    def __init__(self, params: Params):
        assert params.num_limbs == 3
        self.p = params
        self.A, self.Lc, self.F = params.num_advice, params.L, params.num_fixed
        self.n, self.u = params.n, params.usable
        self.lb, self.LB = params.lookup_bits, params.limb_bits
        self.top_bits = 256 - 2 * self.LB
        self.q_top_bits = Q_OFFSET_BITS + 1 - 2 * self.LB
        self.carry_limbs = -(-(self.LB + 6) // self.lb)      # carries stay below 12 * 2^(limb_bits) < 2^(limb_bits + 4)
        self.carry_bits = self.carry_limbs * self.lb
        self.rows = [0] * self.A
        self.advice = [[0] * self.u for _ in range(self.A + self.Lc)]
        self.q_enable = [[0] * self.n for _ in range(self.A)]
        self.q_lookup = [0] * self.n
        self.copies = []               # (cell, cell)
        self.const_copies = []         # (cell, constant index)
        self.constants = {}            # value -> index
        self.lookups = []              # cells (non-selector mode)
        self.stats = {}

    # -- allocation -----------------------------------------------------------------------------------
    def val(self, cell):
        return self.advice[cell // self.n][cell % self.n]

    def assign(self, items, gates):
        """items: ('w', value) new witness | ('c', constant) | ('x', cell) copy of an existing cell.
        Placed consecutively in the gate column with the fewest used rows; gates = offsets with q = 1."""
        c = min(range(self.A), key=lambda i: self.rows[i])
        r0 = self.rows[c]
        if r0 + len(items) > self.u:
            raise DoesNotFit(f"gate column {c} is full ({r0} + {len(items)} > {self.u})")
        out = []
        col = self.advice[c]
        for i, (kind, v) in enumerate(items):
            cell = c * self.n + r0 + i
            if kind == "w":
                col[r0 + i] = v % R
            elif kind == "c":
                col[r0 + i] = v % R
                self.const_copies.append((cell, self.constants.setdefault(v % R, len(self.constants))))
            else:
                col[r0 + i] = self.val(v)
                self.copies.append((v, cell))
            out.append(cell)
        for g in gates:
            self.q_enable[c][r0 + g] = 1
        self.rows[c] = r0 + len(items)
        return out

    def lookup(self, cell):
        if self.p.selector_mode:
            assert cell // self.n == 0
            self.q_lookup[cell % self.n] = 1
        else:
            self.lookups.append(cell)

    def equal(self, a, b):
        self.copies.append((a, b))

    # -- chains: acc' = acc + x*y, cells [init, x0, y0, acc0, x1, y1, acc1, ...] ------------------------------
    def chain(self, init, terms):
        """init: ('c', v) | ('x', cell) | ('w', v);  terms: list of (item, item).  Returns the cell of the final
        accumulator (the init cell when there are no terms)."""
        items = [init]
        acc = init[1] % R if init[0] != "x" else self.val(init[1])
        for x, y in terms:
            xv = x[1] if x[0] != "x" else self.val(x[1])
            yv = y[1] if y[0] != "x" else self.val(y[1])
            acc = (acc + xv * yv) % R
            items += [x, y, ("w", acc)]
        cells = self.assign(items, [3 * i for i in range(len(terms))])
        return cells[-1], cells

    # -- range chip -----------------------------------------------------------------------------------
    def range_limbs(self, value, bits):
        """Witness `value` < 2^bits as a fresh cell that is range-checked; returns the cell."""
        assert 0 <= value < (1 << bits), "range witness out of range"
        lb = self.lb
        k = -(-bits // lb)
        rem = bits % lb
        vs = [(value >> (lb * i)) & ((1 << lb) - 1) for i in range(k)]
        if k == 1:
            (cell,) = self.assign([("w", value)], [])
            limb_cells = [cell]
        else:
            items = [("w", vs[0])]
            acc = vs[0]
            for i in range(1, k):
                acc += vs[i] << (lb * i)
                items += [("w", vs[i]), ("c", 1 << (lb * i)), ("w", acc)]
            cells = self.assign(items, [3 * i for i in range(k - 1)])
            limb_cells = [cells[0]] + [cells[3 * i - 2] for i in range(1, k)]
            cell = cells[-1]
        for lc in limb_cells:
            self.lookup(lc)
        if rem:
            sh = self.assign([("c", 0), ("x", limb_cells[-1]), ("c", 1 << (lb - rem)), ("w", vs[-1] << (lb - rem))], [0])
            self.lookup(sh[3])
        return cell

    def assert_bit(self, cell):
        self.assign([("c", 0), ("x", cell), ("x", cell), ("x", cell)], [0])

    # -- CRT integers ---------------------------------------------------------------------------------
    def split(self, v):
        m = (1 << self.LB) - 1
        return [v & m, (v >> self.LB) & m, v >> (2 * self.LB)]

    def native_of(self, limb_cells):
        c, _ = self.chain(("x", limb_cells[0]), [(("x", limb_cells[1]), ("c", 1 << self.LB)), (("x", limb_cells[2]), ("c", 1 << (2 * self.LB)))])
        return c

    def new_elem(self, value, top_bits=None):
        """Fresh proper CRT integer: limbs range-checked to (limb_bits, limb_bits, top_bits)."""
        lv = self.split(value)
        bits = [self.LB, self.LB, self.top_bits if top_bits is None else top_bits]
        limbs = [self.range_limbs(lv[i], bits[i]) for i in range(3)]
        return Elem(limbs, self.native_of(limbs), value, lv)

    def const_elem(self, value):
        lv = self.split(value)
        cells = self.assign([("c", lv[0]), ("c", lv[1]), ("c", lv[2]), ("c", value % R)], [])
        return Elem(cells[:3], cells[3], value, lv)

    def constrain(self, modulus, pos, neg, k0=0):
        """sum_{(X,Y) in pos} X*Y - sum_{(X,Y) in neg} X*Y + k0 = 0 (mod modulus), X an Elem, Y an Elem or a small
        positive integer.  Proves the integer identity  lhs + 2^258*modulus = q' * modulus  limb by limb modulo
        2^(3*limb_bits) with offset carries, and natively modulo the BN254 scalar field (CRT)."""
        LB = self.LB

        def tv(t):
            X, Y = t
            return X.value * (Y.value if isinstance(Y, Elem) else Y)

        V = sum(tv(t) for t in pos) - sum(tv(t) for t in neg) + k0
        q, rem = divmod(V, modulus)
        Q0 = 1 << Q_OFFSET_BITS
        qp = q + Q0
        bad = rem != 0 or not (0 <= qp < 2 * Q0)
        if bad:                                  # unsatisfiable input: keep laying out cells, the checker will object
            qp = qp % (2 * Q0)
        qe = self.new_elem(qp, self.q_top_bits)
        me = self.split(modulus)
        kpos = (k0 if k0 > 0 else 0) + Q0 * modulus
        kneg = -k0 if k0 < 0 else 0
        kp, kn = self.split(kpos % (1 << (3 * LB))), self.split(kneg)
        OFF = 1 << (self.carry_bits - 1)

        def limb_terms(terms, i):
            out = []
            for X, Y in terms:
                if isinstance(Y, Elem):
                    out += [(("x", X.limbs[j]), ("x", Y.limbs[i - j])) for j in range(i + 1)]
                else:
                    out.append((("x", X.limbs[i]), ("c", Y)))
            return out

        def limb_sum(terms, i):
            s = 0
            for X, Y in terms:
                if isinstance(Y, Elem):
                    s += sum(X.lv[j] * Y.lv[i - j] for j in range(i + 1))
                else:
                    s += X.lv[i] * Y
            return s

        carry_prev_cell = None
        for i in range(3):
            # left chain first (its final value is known once laid out), then the carry cell, then the right chain
            lt = limb_terms(pos, i)
            if i > 0:
                lt.append((("x", carry_prev_cell), ("c", 1)))
            lend, _ = self.chain(("c", kp[i] + (OFF << LB)), lt)
            rinit = kn[i] + (OFF if i > 0 else 0)
            rpart = rinit + limb_sum(neg, i) + sum(qe.lv[j] * me[i - j] for j in range(i + 1))
            diff = self.val(lend) - rpart            # = (carry + OFF) * 2^limb_bits when the relation holds
            assert bad or (diff >= 0 and diff % (1 << LB) == 0 and (diff >> LB) < (1 << self.carry_bits)), "carry range"
            cw = (diff >> LB) % (1 << self.carry_bits)
            carry_cell = self.range_limbs(cw, self.carry_bits)
            rt = limb_terms(neg, i) + [(("x", qe.limbs[j]), ("c", me[i - j])) for j in range(i + 1) if me[i - j]]
            rt.append((("x", carry_cell), ("c", 1 << LB)))
            rend, _ = self.chain(("c", rinit), rt)
            self.equal(lend, rend)
            carry_prev_cell = carry_cell

        def nat_terms(terms):
            return [(("x", X.native), ("x", Y.native) if isinstance(Y, Elem) else ("c", Y)) for X, Y in terms]

        lend, _ = self.chain(("c", kpos % R), nat_terms(pos))
        rend, _ = self.chain(("c", kneg % R), nat_terms(neg) + [(("x", qe.native), ("c", modulus % R))])
        self.equal(lend, rend)

    # -- field ops over a modulus ---------------------------------------------------------------------------
    def assert_less_than(self, a: Elem, bound: int):
        """a < bound (bound < 2^256) as integers: witness d = bound - 1 - a as a proper element, a + d = bound - 1
        limb by limb with carry bits."""
        d = (bound - 1 - a.value) % (1 << 256)
        de = self.new_elem(d)
        b = self.split(bound - 1)
        cin_cell, cin = None, 0
        for i in range(3):
            s = a.lv[i] + de.lv[i] + cin
            cout = (s - b[i]) >> self.LB if i < 2 else 0
            lt = [(("x", de.limbs[i]), ("c", 1))]
            if i > 0:
                lt.append((("x", cin_cell), ("c", 1)))
            lend, _ = self.chain(("x", a.limbs[i]), lt)
            if i < 2:
                (cc,) = self.assign([("w", cout % R)], [])
                self.assert_bit(cc)
                rend, _ = self.chain(("c", b[i]), [(("x", cc), ("c", 1 << self.LB))])
                cin_cell, cin = cc, cout
            else:
                rend, _ = self.chain(("c", b[i]), [])
            self.equal(lend, rend)

    def assert_nonzero(self, a: Elem):
        s, _ = self.chain(("x", a.limbs[0]), [(("x", a.limbs[1]), ("c", 1)), (("x", a.limbs[2]), ("c", 1))])
        sv = self.val(s)
        inv = pow(sv, -1, R) if sv else 0
        self.assign([("c", 0), ("x", s), ("w", inv), ("c", 1)], [0])

    def divide(self, a: Elem, b: Elem, modulus: int) -> Elem:
        inv = pow(b.value, -1, modulus) if b.value % modulus else 0
        u = self.new_elem(a.value * inv % modulus)
        self.constrain(modulus, [(u, b)], [(a, 1)])
        return u

    # -- bits and window indicators -------------------------------------------------------------------------
    def to_bits(self, cell, nbits):
        v = self.val(cell)
        bits = [(v >> i) & 1 for i in range(nbits)]
        items = [("w", bits[0])]
        acc = bits[0]
        for i in range(1, nbits):
            acc += bits[i] << i
            items += [("w", bits[i]), ("c", 1 << i), ("w", acc)]
        cells = self.assign(items, [3 * i for i in range(nbits - 1)])
        self.equal(cells[-1], cell)
        bcells = [cells[0]] + [cells[3 * i - 2] for i in range(1, nbits)]
        for b in bcells:
            self.assert_bit(b)
        return bcells

    def indicator(self, bits_hi_to_lo):
        """[b3, b2, b1, b0] -> 16 cells, cell j = 1 iff the window's value is j."""
        b = bits_hi_to_lo[0]
        c = self.assign([("w", 1 - self.val(b)), ("x", b), ("c", 1), ("c", 1)], [0])
        ind = [c[0], b]
        for b in bits_hi_to_lo[1:]:
            bv = self.val(b)
            nxt = []
            for e in ind:
                ev = self.val(e)
                m = self.assign([("c", 0), ("x", e), ("x", b), ("w", ev * bv)], [0])
                s = self.assign([("w", ev - ev * bv), ("x", m[3]), ("c", 1), ("x", e)], [0])
                nxt += [s[0], m[3]]
            ind = nxt
        return ind

    def select_elem(self, ind, table, const: bool):
        """sum_j ind_j * table_j: limbs by inner products; table entries are Elems (cells) or integers (constants)."""
        sel = [self.val(i) for i in ind].index(1)
        limbs = []
        for i in range(3):
            if const:
                terms = [(("x", ind[j]), ("c", self.split(table[j])[i])) for j in range(len(ind))]
            else:
                terms = [(("x", ind[j]), ("x", table[j].limbs[i])) for j in range(len(ind))]
            c, _ = self.chain(("c", 0), terms)
            limbs.append(c)
        value = table[sel] if const else table[sel].value
        return Elem(limbs, self.native_of(limbs), value, self.split(value))

    # -- curve ops (P-256: y^2 = x^3 - 3x + b) ---------------------------------------------------------------
    def ec_add(self, P, Q, strict=False):
        (x1, y1), (x2, y2) = P, Q
        p = P256_P
        dx = (x2.value - x1.value) % p
        if strict:                                    # x1 != x2 (mod p): (x2 - x1) has an inverse
            t = self.new_elem(pow(dx, -1, p) if dx else 0)
            self.constrain(p, [(t, x2)], [(t, x1)], k0=-1)
        lam_v = (y2.value - y1.value) * pow(dx, -1, p) % p if dx else 0
        lam = self.new_elem(lam_v)
        self.constrain(p, [(lam, x2), (y1, 1)], [(lam, x1), (y2, 1)])
        x3 = self.new_elem((lam_v * lam_v - x1.value - x2.value) % p)
        self.constrain(p, [(lam, lam)], [(x1, 1), (x2, 1), (x3, 1)])
        y3 = self.new_elem((lam_v * (x1.value - x3.value) - y1.value) % p)
        self.constrain(p, [(lam, x1)], [(lam, x3), (y1, 1), (y3, 1)])
        return x3, y3

    def ec_double(self, P):
        x, y = P
        p = P256_P
        lam_v = (3 * x.value * x.value - 3) * pow(2 * y.value, -1, p) % p if y.value % p else 0
        lam = self.new_elem(lam_v)
        self.constrain(p, [(lam, y), (lam, y)], [(x, x), (x, x), (x, x)], k0=3)
        x3 = self.new_elem((lam_v * lam_v - 2 * x.value) % p)
        self.constrain(p, [(lam, lam)], [(x, 2), (x3, 1)])
        y3 = self.new_elem((lam_v * (x.value - x3.value) - y.value) % p)
        self.constrain(p, [(lam, x)], [(lam, x3), (y, 1), (y3, 1)])
        return x3, y3

    def assert_on_curve(self, P):
        x, y = P
        p = P256_P
        x2 = self.new_elem(x.value * x.value % p)
        self.constrain(p, [(x, x)], [(x2, 1)])
        self.constrain(p, [(y, y), (x, 3)], [(x2, x)], k0=-P256_B)

    def scalar_windows(self, u: Elem):
        """The 3 * limb_bits bits of u as 4-bit window indicators, most significant window first."""
        assert self.LB % WINDOW == 0 or True
        bits = []
        for i in range(3):
            bits += self.to_bits(u.limbs[i], self.LB)
        while len(bits) % WINDOW:
            (z,) = self.assign([("c", 0)], [])
            bits.append(z)
        nw = len(bits) // WINDOW
        return [self.indicator([bits[WINDOW * w + j] for j in range(WINDOW - 1, -1, -1)]) for w in range(nw - 1, -1, -1)]


def num_windows(limb_bits):
    return -(-3 * limb_bits // WINDOW)


def var_offset_scalar(limb_bits):
    """The variable-base accumulator ends at u*PK + c*B2 with c = sum_w 16^w."""
    return sum(1 << (WINDOW * w) for w in range(num_windows(limb_bits))) % P256_N


def fixed_tables(limb_bits):
    """T[w][j] = j * 16^w * G + o_w * B3 with o_w = 2^w for w < W-1 and o_{W-1} = -(2^(W-1) - 1): the offsets sum
    to zero.  Returns the W x 16 table of affine points."""
    W = num_windows(limb_bits)
    tabs = []
    base = P256_G
    for w in range(W):
        o = (1 << w) if w < W - 1 else (-((1 << (W - 1)) - 1)) % P256_N
        off = ec_mul(OFFSET_FIX, o)
        row = [off]
        for j in range(1, 16):
            row.append(ec_add(row[-1], base))
        tabs.append(row)
        for _ in range(WINDOW):
            base = ec_add(base, base)
    return tabs


_FIXED_CACHE = {}


def synthesize(params: Params, pubkey, r: int, s: int, msghash: int) -> Builder:
    """Lays out the whole circuit for one assertion and returns the Builder (advice, selectors, copies, lookups,
    constants).  Never raises on an invalid signature: the assignment it returns then violates constraints."""
    b = Builder(params)
    n, p = P256_N, P256_P
    # ecdsa_p256.rs:139-177: load m, r, s (scalar field) and the public key (base field)
    m_e, r_e, s_e = b.new_elem(msghash), b.new_elem(r), b.new_elem(s)
    pk = (b.new_elem(pubkey[0]), b.new_elem(pubkey[1]))
    b.assert_on_curve(pk)
    # r, s in [1, n-1]
    for e in (r_e, s_e):
        b.assert_nonzero(e)
        b.assert_less_than(e, n)
    u1 = b.divide(m_e, s_e, n)
    u2 = b.divide(r_e, s_e, n)
    b.assert_less_than(u1, n)
    b.assert_less_than(u2, n)

    # variable base: table T[d] = d*PK + B2, accumulator acc = 16*acc + T[window]
    B2 = (b.const_elem(OFFSET_VAR[0]), b.const_elem(OFFSET_VAR[1]))
    table = [B2]
    for _ in range(15):
        table.append(b.ec_add(table[-1], pk))
    tx, ty = [t[0] for t in table], [t[1] for t in table]
    # everything the two window loops consume is laid out before them (the product synthesises the loops' windows on
    # several host threads, each starting from a recorded row offset): window indicators of both scalars and the
    # fixed-base accumulator's constant start point -(c*B2), which cancels the variable part's offset in the final addition
    var_windows = b.scalar_windows(u2)
    fix_windows = b.scalar_windows(u1)
    key = params.limb_bits
    if key not in _FIXED_CACHE:
        _FIXED_CACHE[key] = (fixed_tables(key), ec_neg(ec_mul(OFFSET_VAR, var_offset_scalar(key))))
    tabs, start = _FIXED_CACHE[key]
    facc = (b.const_elem(start[0]), b.const_elem(start[1]))
    acc = None
    for ind in var_windows:
        sel = (b.select_elem(ind, tx, False), b.select_elem(ind, ty, False))
        if acc is None:
            acc = sel
        else:
            for _ in range(WINDOW):
                acc = b.ec_double(acc)
            acc = b.ec_add(acc, sel)
    W = len(tabs)
    for wi, ind in enumerate(fix_windows):
        w = W - 1 - wi
        sel = (b.select_elem(ind, [t[0] for t in tabs[w]], True), b.select_elem(ind, [t[1] for t in tabs[w]], True))
        facc = b.ec_add(facc, sel)
    # R = u1*G + u2*PK, strict addition; R.x == r as integers (limb by limb)
    Rx, _ = b.ec_add(acc, facc, strict=True)
    for i in range(3):
        b.equal(Rx.limbs[i], r_e.limbs[i])
    finalize(b)
    return b


def finalize(b: Builder):
    """Copies the cells to look up into the lookup advice columns (non-selector mode): cell i -> column i mod L,
    row i div L — the analogue of `fp_chip.finalize(ctx)` (ecdsa_p256.rs:193-195)."""
    if b.p.selector_mode:
        return
    L = b.Lc
    if len(b.lookups) > L * b.u:
        raise DoesNotFit(f"{len(b.lookups)} lookup cells > {L} x {b.u}")
    for i, cell in enumerate(b.lookups):
        col, row = b.A + i % L, i // L
        b.advice[col][row] = b.val(cell)
        b.copies.append((cell, col * b.n + row))


# ---- structure for keygen ------------------------------------------------------------------------------------
def fixed_columns(b: Builder):
    """[constants (F columns), table, q_enable (A columns), (q_lookup)] as lists of n canonical integers."""
    n, F = b.n, b.F
    consts = [[0] * n for _ in range(F)]
    for v, idx in b.constants.items():
        if idx // F >= b.u:
            raise DoesNotFit("too many constants")
        consts[idx % F][idx // F] = v
    table = [0] * n
    for i in range(min(1 << b.lb, b.u)):
        table[i] = i
    cols = consts + [table] + [list(q) for q in b.q_enable]
    if b.p.selector_mode:
        cols.append(list(b.q_lookup))
    return cols


def permutation_mapping(b: Builder):
    """Per permutation column [constants.., gate advice.., lookup advice..] a list of n (col', row') successors: the
    cells of each copy class in increasing (col, row) order, closed into a cycle."""
    n, F = b.n, b.F
    ncols = F + b.A + b.Lc
    parent = {}

    def find(x):
        while parent.setdefault(x, x) != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    def union(x, y):
        rx, ry = find(x), find(y)
        if rx != ry:
            parent[max(rx, ry)] = min(rx, ry)

    for a, c in b.copies:
        union(F * n + a, F * n + c)
    for cell, idx in b.const_copies:
        union(F * n + cell, (idx % F) * n + idx // F)
    classes = {}
    for x in sorted(parent):
        classes.setdefault(find(x), []).append(x)
    mapping = [[(c, r) for r in range(n)] for c in range(ncols)]
    for members in classes.values():
        for i, x in enumerate(members):
            y = members[(i + 1) % len(members)]
            mapping[x // n][x % n] = (y // n, y % n)
    return mapping


def check(b: Builder, fixed=None, mapping=None):
    """MockProver::verify analogue: returns a list of violated constraints (empty = satisfied)."""
    errs = []
    n, u, A, F = b.n, b.u, b.A, b.F
    fixed = fixed or fixed_columns(b)
    mapping = mapping or permutation_mapping(b)
    for c in range(A):
        q = fixed[F + 1 + c]
        col = b.advice[c]
        for i in range(n):
            if q[i]:
                if i + 3 >= u:
                    errs.append(("gate-out-of-rows", c, i))
                elif (col[i] + col[i + 1] * col[i + 2] - col[i + 3]) % R:
                    errs.append(("gate", c, i))
                    if len(errs) > 20:
                        return errs
    T = min(1 << b.lb, u)
    if b.p.selector_mode:
        ql = fixed[F + 1 + A]
        for i in range(u):
            if ql[i] and not (0 <= b.advice[0][i] < T):
                errs.append(("lookup", 0, i))
    else:
        for l in range(b.Lc):
            for i, v in enumerate(b.advice[A + l]):
                if not (0 <= v < T):
                    errs.append(("lookup", l, i))

    def value(c, r):
        return fixed[c][r] if c < F else b.advice[c - F][r]

    for c, colmap in enumerate(mapping):
        for r, (cc, rr) in enumerate(colmap):
            if (cc, rr) != (c, r):
                if r >= u or rr >= u or value(c, r) != value(cc, rr):
                    errs.append(("copy", c, r, cc, rr))
                    if len(errs) > 20:
                        return errs
    return errs


def cell_counts(b: Builder):
    return {"gate_rows": list(b.rows), "gate_cells": sum(b.rows), "lookups": sum(b.q_lookup) if b.p.selector_mode else len(b.lookups),
            "constants": len(b.constants), "copies": len(b.copies) + len(b.const_copies)}
