"""Host logic of the product's circuit description (webauthn-halo2_b200/circuit.py), checked on the CPU
with the oracle's MockProver-style checker: the synthetic ECDSA-shaped assignment satisfies every gate,
lookup and copy constraint for each of the reference's nine configs' column layouts (scaled down in k)."""
import hashlib

import numpy as np
import pytest

from oracle import halo2_ref as h, synth_circuit as sc


@pytest.mark.parametrize("degree,A,L,F,lb", [(5, 1, 1, 1, 4), (6, 2, 1, 1, 5), (6, 4, 1, 1, 5), (7, 8, 2, 1, 6), (6, 17, 3, 1, 5), (6, 5, 2, 2, 4), (7, 3, 1, 4, 6)])
def test_synthetic_assignment_satisfies_the_constraint_system(zkw, degree, A, L, F, lb):
    params = zkw.CircuitParams("Simple", degree, A, L, F, lb, 88, 3)
    circ = zkw.SyntheticEcdsaCircuit(params)
    oshape = h.Shape(degree, circ.A, circ.L, circ.F)
    assert (circ.nfixed, circ.nperm) == (oshape.num_fixed_cols, len(oshape.perm_columns()))
    fixed = [[int(x) for x in c] for c in circ.fixed_columns()]
    mapping = [[(int(a), int(b)) for a, b in m] for m in circ.permutation_mapping()]
    for assertion in (b"a", b"b"):
        advice = [[int(x) for x in c] for c in circ.synthesize(assertion)]
        assert all(len(c) <= circ.u for c in advice)
        assert sc.check_satisfied(oshape, fixed, mapping, advice)
    assert any((x != y).any() for x, y in zip(circ.synthesize(b"a"), circ.synthesize(b"b")))   # keyed by the assertion
    assert all((x == y).all() for x, y in zip(circ.synthesize(b"a"), circ.synthesize(b"a")))   # and deterministic


def test_permutation_mapping_is_a_permutation(zkw):
    circ = zkw.SyntheticEcdsaCircuit(zkw.CircuitParams("Simple", 7, 4, 1, 1, 6, 88, 3))
    maps = circ.permutation_mapping()
    cells = set()
    for m in maps:
        cells.update((int(a), int(b)) for a, b in m)
    assert len(cells) == circ.nperm * circ.n           # bijection on cells
    # every copy pair is a 2-cycle
    for ca, ra, cb, rb in circ.copy_pairs():
        for i in range(len(ca)):
            assert tuple(maps[ca[i]][ra[i]]) == (cb[i], rb[i]) and tuple(maps[cb[i]][rb[i]]) == (ca[i], ra[i])


def test_reference_configs_table(zkw):
    """halo2-circuits/src/configs/bench_ecdsa.config, all nine lines."""
    assert zkw.CircuitParams.for_degree(19).__dict__ == dict(strategy="Simple", degree=19, num_advice=1, num_lookup_advice=1, num_fixed=1,
                                                             lookup_bits=18, limb_bits=88, num_limbs=3)
    assert zkw.CircuitParams.for_degree(11).num_advice == 291 and zkw.CircuitParams.for_degree(11).num_fixed == 4
    p = zkw.CircuitParams.from_json('{"strategy":"Simple","degree":17,"num_advice":4,"num_lookup_advice":1,"num_fixed":1,"lookup_bits":16,"limb_bits":88,"num_limbs":3}')
    assert p == zkw.CircuitParams.for_degree(17)
    with pytest.raises(ValueError):
        zkw.CircuitParams.for_degree(20)


def test_validate_assertion_mirrors_reference_unwraps(zkw):
    gx = bytes.fromhex("6b17d1f2e12c4247f8bce6e563a440f277037d812deb33a0f4a13945d898c296")[::-1]
    gy = bytes.fromhex("4fe342e2fe1a7f9b8ee7eb4a7c0f9e162bce33576b315ececbb6406837bf51f5")[::-1]
    ok = hashlib.sha256(b"x").digest()[:31] + b"\0"
    zkw.validate_assertion(gx, gy, ok, ok, ok)
    for bad in [(gx[:-1], gy, ok, ok, ok), (gx, gy, b"\xff" * 32, ok, ok), (gx, gy, ok, ok, b"\xff" * 32), (gy, gx, ok, ok, ok),
                (b"\xff" * 32, gy, ok, ok, ok)]:
        with pytest.raises(ValueError):
            zkw.validate_assertion(*bad)


@pytest.mark.parametrize("degree", [19, 17, 12, 7])
def test_native_synthesis_matches_numpy_statement(zkw, degree):
    """zkw_synth_witness (host C++ in the library, the routine inside the timed end-to-end path) and
    SyntheticEcdsaCircuit.synthesize (numpy) fill the same cells."""
    params = zkw.CircuitParams.for_degree(degree) if degree != 7 else zkw.CircuitParams("Simple", 7, 4, 1, 1, 6, 88, 3)
    circ = zkw.SyntheticEcdsaCircuit(params)
    shape = zkw.CircuitShape.from_config(params.degree, params.num_advice, params.num_lookup_advice, params.num_fixed)
    for assertion in (b"", b"assertion-1", bytes(range(160))):
        want = circ.synthesize(assertion)
        got = zkw.native.synth_witness(shape, params.lookup_bits, assertion)
        assert len(want) == len(got)
        for w, g in zip(want, got):
            assert w.shape == g.shape and (w == g).all()
