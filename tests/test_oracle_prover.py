"""Oracle-side create_proof / verify_proof round trips on synthetic ECDSA-shaped circuits (CPU)."""
import numpy as np
import pytest

from oracle import cpu, halo2_ref as h, pyref as pr, synth_circuit as sc

TAU = 0x1234567890ABCDEF1234567890ABCDEF % pr.R


def dev_srs(shape):
    n = shape.n
    dom = shape.domain()
    g = cpu.srs_powers(n, cpu.fr_to_mont([TAU])[0])
    c = (pow(TAU, n, pr.R) - 1) * pow(n, -1, pr.R) % pr.R
    ls = [c * pow(dom.omega, i, pr.R) % pr.R * pow((TAU - pow(dom.omega, i, pr.R)) % pr.R, -1, pr.R) % pr.R for i in range(n)]
    return g, cpu.g1_fixed_base_mul(cpu.fr_to_mont(ls))


@pytest.mark.parametrize("k,A,L,F,kind", [(5, 1, 0, 1, "evm"), (5, 1, 0, 1, "blake2b"), (6, 4, 1, 1, "evm"), (5, 2, 1, 2, "blake2b")])
def test_round_trip(k, A, L, F, kind):
    shape = h.Shape(k, A, L, F)
    fixed, mapping, advice = sc.build(shape, seed=k)
    assert sc.check_satisfied(shape, fixed, mapping, advice)
    g, gl = dev_srs(shape)
    pk = h.keygen(shape, gl, fixed, h.sigma_from_cycles(shape, mapping))
    proof = h.create_proof(pk, g, gl, advice, seed=42, kind=kind)
    assert h.verify_proof(pk.vk, proof, kind, tau=TAU)
    # same inputs + same blinding stream => same bytes; another seed => another proof that still verifies
    assert proof == h.create_proof(pk, g, gl, advice, seed=42, kind=kind)
    other = h.create_proof(pk, g, gl, advice, seed=43, kind=kind)
    assert other != proof and h.verify_proof(pk.vk, other, kind, tau=TAU)
    bad = bytearray(proof)
    bad[len(bad) // 2] ^= 1
    assert not h.verify_proof(pk.vk, bytes(bad), kind, tau=TAU)
    wrong = [list(c) for c in advice]
    wrong[0][3] = (wrong[0][3] + 1) % pr.R
    assert not h.verify_proof(pk.vk, h.create_proof(pk, g, gl, wrong, seed=42, kind=kind), kind, tau=TAU)


@pytest.mark.parametrize("k,A,L,F,kind", [(5, 1, 0, 1, "blake2b"), (6, 4, 1, 1, "blake2b"), (5, 2, 1, 2, "evm")])
def test_round_trip_shplonk(k, A, L, F, kind):
    shape = h.Shape(k, A, L, F)
    fixed, mapping, advice = sc.build(shape, seed=k)
    g, gl = dev_srs(shape)
    pk = h.keygen(shape, gl, fixed, h.sigma_from_cycles(shape, mapping))
    proof = h.create_proof(pk, g, gl, advice, seed=42, kind=kind, multiopen="shplonk")
    assert h.verify_proof(pk.vk, proof, kind, tau=TAU, multiopen="shplonk")
    assert not h.verify_proof(pk.vk, proof, kind, tau=TAU, multiopen="gwc")
    for pos in (10, len(proof) // 2, len(proof) - 40, len(proof) - 1):
        bad = bytearray(proof)
        bad[pos] ^= 1
        assert not h.verify_proof(pk.vk, bytes(bad), kind, tau=TAU, multiopen="shplonk")
    if (k, A, L, F, kind) == (5, 1, 0, 1, "blake2b"):
        assert len(proof) == 960      # the k = 19 layout: halo2-circuits/src/results/ecdsa_bench.csv:2
    if (k, A, L, F, kind) == (6, 4, 1, 1, "blake2b"):
        assert len(proof) == 1920     # the k = 17 layout: ecdsa_bench.csv:4


def test_lagrange_interpolate():
    pts = [3, 7, 11, 20]
    coeffs = [5, 0, 9, 2]
    evals = [pr.poly_eval(coeffs, p) for p in pts]
    assert h.lagrange_interpolate(pts, evals) == coeffs


def test_proof_sizes_match_reference():
    """k=17 config under the EVM transcript: 2720 bytes = the golden proof's length
    (contracts/test/P256Account.t.sol:120).  k=19 config under Blake2b: 10 + 5 (GWC) commitments and 18
    evaluations; the reference's csv (SHPLONK, 2 opening points instead of 5) is 3*32 bytes shorter: 960."""
    s17 = h.Shape(17, 4, 1, 1)
    pts = s17.num_advice_cols + 3 * s17.num_lookups + s17.num_perm_sets + 1 + s17.quotient_pieces + 6
    evs = len(s17.advice_queries()) + len(s17.fixed_queries()) + 1 + len(s17.perm_columns()) + 3 * s17.num_perm_sets - 1 + 5
    assert 64 * pts + 32 * evs == 2720
    s19 = h.Shape(19, 1, 0, 1)
    pts = s19.num_advice_cols + 3 * s19.num_lookups + s19.num_perm_sets + 1 + s19.quotient_pieces + 5
    evs = len(s19.advice_queries()) + len(s19.fixed_queries()) + 1 + len(s19.perm_columns()) + 3 * s19.num_perm_sets - 1 + 5
    assert (pts, evs) == (15, 18) and 32 * (pts - 3) + 32 * evs == 960


def test_lookup_permutation_properties():
    shape = h.Shape(5, 1, 0, 1)
    fixed, mapping, advice = sc.build(shape, seed=3)
    adv = [list(advice[0]) + [0] * (shape.usable_rows - len(advice[0]))]
    inp = h.lookup_input_values(shape, [adv[0] + [0] * 7], fixed, 0)
    a, s = h.permute_expression_pair(shape, inp, fixed[shape.table_col], 1, 0)
    u = shape.usable_rows
    assert sorted(a[:u]) == a[:u] == sorted(inp[:u])
    assert sorted(s[:u]) == sorted(fixed[shape.table_col][:u])
    assert a[0] == s[0] and all(a[i] == s[i] or a[i] == a[i - 1] for i in range(1, u))
    with pytest.raises(ValueError):
        bad = list(inp)
        bad[2] = 12345678
        h.permute_expression_pair(shape, bad, fixed[shape.table_col], 1, 0)
