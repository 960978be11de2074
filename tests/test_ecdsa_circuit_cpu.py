"""The real P-256 ECDSA verification circuit on the CPU (no GPU needed: layout and witness synthesis are host code).

  * oracle/ecdsa_circuit.py (plain Python integers) lays out the circuit for a valid signature and its
    MockProver-style checker finds every gate, lookup and copy constraint satisfied — the reference's own circuit
    test (halo2-circuits/src/ecc/ecdsa_p256.rs:209-248: random key, random message, MockProver::verify == Ok) —
    for the reference's benchmark configs (halo2-circuits/src/configs/bench_ecdsa.config), all of which it fits;
  * an invalid signature has NO satisfying assignment (the honest assignment violates the R.x == r copies);
  * the product's C++ synthesis (webauthn-halo2_b200/csrc/ecdsa_circuit.cpp, through the C ABI) agrees with the
    oracle cell for cell: advice columns, fixed columns (constants, table, selectors) and the permutation.
"""
import numpy as np
import pytest

from oracle import ecdsa_circuit as ec
from tests.assertions import signed_assertion, signed_ints, verify_ints, N as P256_N

CONFIGS = {19: (1, 1, 1, 18, 88), 18: (2, 1, 1, 17, 88), 17: (4, 1, 1, 16, 88), 16: (8, 2, 1, 15, 90), 15: (17, 3, 1, 14, 90),
           14: (34, 6, 1, 13, 91), 13: (68, 12, 1, 12, 88), 12: (139, 24, 2, 11, 88), 11: (291, 53, 4, 10, 88)}


def _params(k):
    A, L, F, lb, limb = CONFIGS[k]
    return ec.Params(k, A, L, F, lb, limb)


def _synth(k, a):
    return ec.synthesize(_params(k), (a["pubkey_x"], a["pubkey_y"]), a["r"], a["s"], a["msg_hash"])


def _ints(col):
    a = col.astype(object)
    return list(a[:, 0] + (a[:, 1] << 64) + (a[:, 2] << 128) + (a[:, 3] << 192))


@pytest.mark.parametrize("k", [19, 17, 16, 12])
def test_oracle_circuit_is_satisfied_by_a_valid_signature(k):
    a = signed_ints(k)
    assert verify_ints(a)
    b = _synth(k, a)
    assert ec.check(b) == []
    counts = ec.cell_counts(b)
    assert max(counts["gate_rows"]) <= b.u and counts["lookups"] <= max(1, b.Lc) * b.u


def test_every_reference_config_fits():
    """bench_ecdsa.config lines 1-9: the layout fits the usable rows of every column set the reference benchmarks."""
    a = signed_ints(1)
    for k in CONFIGS:
        b = _synth(k, a)      # raises DoesNotFit otherwise
        assert max(b.rows) <= b.u


@pytest.mark.parametrize("what", ["msg", "r", "s", "pubkey"])
def test_invalid_signature_has_no_satisfying_assignment(what):
    a = signed_ints(5)
    if what == "msg":
        a["msg_hash"] ^= 1
    elif what == "r":
        a["r"] = (a["r"] + 1) % P256_N
    elif what == "s":
        a["s"] = (a["s"] + 1) % P256_N
    else:
        other = signed_ints(6)
        a["pubkey_x"], a["pubkey_y"] = other["pubkey_x"], other["pubkey_y"]
    assert not verify_ints(a)
    errs = ec.check(_synth(19, a))
    assert errs and all(e[0] == "copy" for e in errs)     # every relation holds except R.x == r


def _sign(sk, k, m):
    from tests.assertions import G, mul
    pk = mul(G, sk)
    r = mul(G, k)[0] % P256_N
    s = pow(k, -1, P256_N) * (m + r * sk) % P256_N
    return {"pubkey_x": pk[0], "pubkey_y": pk[1], "r": r, "s": s, "msg_hash": m}


@pytest.mark.parametrize("name,sk,k,m", [
    ("zero message hash (u1 = 0: every fixed-base window is zero)", 0x1234567, 0x7654321, 0),
    ("secret key 1 (PK = G)", 1, 5, 12345),
    ("tiny nonce and message (u2 has long zero runs)", 3, 1, 1),
    ("top-heavy values", P256_N - 2, P256_N - 3, P256_N - 1),
])
def test_edge_case_signatures_are_satisfied(zkw, name, sk, k, m):
    """Valid signatures at the corners of the scalar range: the offset points keep every accumulator away from the identity,
    the oracle's assignment is satisfied, and the product's synthesis agrees with it and reports the signature as valid."""
    a = _sign(sk, k, m)
    assert verify_ints(a), name
    b = _synth(17, a)
    assert ec.check(b) == [], name
    c = zkw.EcdsaCircuit(zkw.CircuitParams.for_degree(17))
    try:
        adv = c.synthesize(*[a[key].to_bytes(32, "little") for key in ("pubkey_x", "pubkey_y", "r", "s", "msg_hash")])
        for i, col in enumerate(adv):
            assert _ints(col) == b.advice[i][: col.shape[0]], (name, i)
    finally:
        c.close()


def test_degenerate_inputs_are_unsatisfiable_not_crashes():
    a = signed_ints(7)
    for key, val in (("s", 0), ("r", 0), ("r", P256_N), ("s", P256_N + 5)):
        bad = dict(a)
        bad[key] = val
        assert ec.check(_synth(17, bad))
    off = dict(a)
    off["pubkey_y"] = (a["pubkey_y"] + 1) % ec.P256_P          # not on the curve
    assert ec.check(_synth(17, off))


@pytest.mark.parametrize("k", [19, 17, 13, 11])
def test_product_synthesis_matches_oracle_cell_for_cell(zkw, k):
    P = zkw.CircuitParams.for_degree(k)
    c = zkw.EcdsaCircuit(P)
    try:
        ai, ab = signed_ints(40 + k), signed_assertion(40 + k)
        b = _synth(k, ai)
        assert c.rows[: b.A] == b.rows
        adv = c.synthesize(ab["pubkey_x"], ab["pubkey_y"], ab["r"], ab["s"], ab["msg_hash"])
        assert len(adv) == b.A + b.Lc
        for i, col in enumerate(adv):
            got = _ints(col)
            assert got == b.advice[i][: len(got)], f"advice column {i}"
            assert not any(b.advice[i][len(got):])
        for i, (col, want) in enumerate(zip(c.fixed_columns(), ec.fixed_columns(b))):
            assert _ints(col) == want, f"fixed column {i}"
        for i, (m, want) in enumerate(zip(c.permutation_mapping(), ec.permutation_mapping(b))):
            assert [tuple(int(y) for y in x) for x in m] == want, f"permutation column {i}"
        st = c.stats()
        oc = ec.cell_counts(b)
        assert (st["gate_cells"], st["lookups"], st["constants"]) == (oc["gate_cells"], oc["lookups"], oc["constants"])
    finally:
        c.close()


@pytest.mark.parametrize("k", [19, 16])
def test_threaded_synthesis_is_identical(zkw, monkeypatch, k):
    """zkw_ecdsa_synthesize splits the window loops over worker threads that start from recorded row offsets and from the
    accumulator values of a projective pre-pass: any thread count writes exactly the same cells."""
    c = zkw.EcdsaCircuit(zkw.CircuitParams.for_degree(k))
    try:
        ab = signed_assertion(70 + k)
        args = [ab["pubkey_x"], ab["pubkey_y"], ab["r"], ab["s"], ab["msg_hash"]]
        monkeypatch.setenv("ZKW_SYNTH_THREADS", "1")
        want = c.synthesize(*args)
        for threads in ("2", "3", "5", "16"):
            monkeypatch.setenv("ZKW_SYNTH_THREADS", threads)
            got = c.synthesize(*args)
            assert all(np.array_equal(a, b) for a, b in zip(got, want)), threads
        bad = list(args)
        bad[4] = bytes([bad[4][0] ^ 1]) + bad[4][1:]
        monkeypatch.setenv("ZKW_SYNTH_THREADS", "1")
        want_bad = c.synthesize(*bad, allow_invalid=True)
        monkeypatch.setenv("ZKW_SYNTH_THREADS", "4")
        with pytest.raises(zkw.InvalidSignature):
            c.synthesize(*bad)
        assert all(np.array_equal(a, b) for a, b in zip(c.synthesize(*bad, allow_invalid=True), want_bad))
    finally:
        c.close()


def test_product_refuses_invalid_signatures(zkw):
    c = zkw.EcdsaCircuit(zkw.CircuitParams.for_degree(17))
    try:
        ab = signed_assertion(9)
        args = [ab["pubkey_x"], ab["pubkey_y"], ab["r"], ab["s"], ab["msg_hash"]]
        c.synthesize(*args)
        for i in (2, 3, 4):
            bad = list(args)
            bad[i] = bytes([bad[i][0] ^ 1]) + bad[i][1:]
            with pytest.raises(zkw.InvalidSignature):
                c.synthesize(*bad)
            # forced through, the assignment equals the oracle's (unsatisfied) one
            ai = signed_ints(9)
            key = ("r", "s", "msg_hash")[i - 2]
            ai[key] ^= 1
            adv = c.synthesize(*bad, allow_invalid=True)
            b = _synth(17, ai)
            assert _ints(adv[0]) == b.advice[0][: adv[0].shape[0]]
        zero = bytes(32)
        for i in (2, 3):
            bad = list(args)
            bad[i] = zero
            with pytest.raises(zkw.InvalidSignature):
                c.synthesize(*bad)
    finally:
        c.close()


def test_offset_points_are_on_the_curve_and_hash_derived():
    import hashlib
    for tag, pt in ((b"zkw-b200 ecdsa variable-base offset", ec.OFFSET_VAR), (b"zkw-b200 ecdsa fixed-base offset", ec.OFFSET_FIX)):
        x, y = pt
        assert (y * y - (x * x * x - 3 * x + ec.P256_B)) % ec.P256_P == 0 and y % 2 == 0
        assert any(int.from_bytes(hashlib.sha256(tag + c.to_bytes(4, "big")).digest(), "big") % ec.P256_P == x for c in range(64))
    tabs = ec.fixed_tables(88)
    assert len(tabs) == 66
    # window offsets cancel: sum_w T[w][0] is the identity
    acc = None
    for row in tabs:
        acc = ec.ec_add(acc, row[0])
    assert acc is None
