"""Committed known-answer vectors (tests/golden/kat_small.json, made by tools/gen_golden_vectors.py):
the oracle must reproduce them on the CPU, the device must reproduce them on the GPU."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "kat_small.json")))
TAU = int(KAT["tau"], 16)


def ints(hexlist):
    return [int(x, 16) for x in hexlist]


def test_oracle_reproduces_fixture(oracle):
    from oracle import halo2_ref as h, pyref as pr
    t = KAT["ntt16"]
    a = oracle.fr_to_mont(ints(t["input"]))
    assert np.array_equal(a, oracle.fr_random(16, t["seed"]))
    assert oracle.fr_from_mont(oracle.best_fft(a, oracle.fr_to_mont([int(t["omega"], 16)])[0])) == ints(t["output"])
    assert ints(t["output"]) == pr.dft_naive(ints(t["input"]), int(t["omega"], 16))
    t = KAT["coset_ext_k3"]
    d = oracle.Domain.new(4, 3)
    assert oracle.fr_from_mont(d.coeff_to_extended(oracle.fr_to_mont(ints(t["coeffs"])))) == ints(t["extended"])
    t = KAT["msm8"]
    s = oracle.fr_random(8, t["scalar_seed"])
    b = oracle.g1_fixed_base_mul(oracle.fr_random(8, t["base_scalar_seed"]))
    assert list(oracle.g1_affine_to_ints(oracle.g1_to_affine(oracle.best_multiexp(s, b))[0])) == ints(t["result_xy"])
    g4 = oracle.srs_powers(4, oracle.fr_to_mont([TAU])[0])
    assert [list(oracle.g1_affine_to_ints(p)) for p in g4] == [ints(p) for p in KAT["srs_g4"]]
    assert ints(KAT["srs_g4"][0]) == [1, 2]


def test_oracle_verifier_accepts_fixture_proofs(oracle):
    from oracle import halo2_ref as h
    t = KAT["proof_k5"]
    vk = h.VerifyingKey(h.Shape(*t["shape"]), [tuple(ints(p)) for p in t["vk_fixed"]], [tuple(ints(p)) for p in t["vk_perm"]],
                        int(t["vk_digest"], 16))
    for key, hx in t["proofs"].items():
        kind, mo = key.split("/")
        assert h.verify_proof(vk, bytes.fromhex(hx), kind, tau=TAU, multiopen=mo), key
    assert len(bytes.fromhex(t["proofs"]["blake2b/shplonk"])) == 960


@pytest.mark.gpu
def test_device_reproduces_fixture(zkw, oracle, ctx):
    t = KAT["ntt16"]
    a = oracle.fr_to_mont(ints(t["input"]))
    assert oracle.fr_from_mont(ctx.ntt(a, oracle.fr_to_mont([int(t["omega"], 16)])[0])) == ints(t["output"])
    t = KAT["coset_ext_k3"]
    assert oracle.fr_from_mont(ctx.coeff_to_extended(oracle.fr_to_mont(ints(t["coeffs"])), 5)) == ints(t["extended"])
    t = KAT["msm8"]
    s = oracle.fr_random(8, t["scalar_seed"])
    b = oracle.g1_fixed_base_mul(oracle.fr_random(8, t["base_scalar_seed"]))
    assert list(oracle.g1_affine_to_ints(ctx.msm(s, b)[:8])) == ints(t["result_xy"])


@pytest.mark.gpu
def test_device_prover_reproduces_fixture_proofs(zkw, oracle):
    """Same SRS, fixed columns, permutation, witness and blinding seed as the fixture: identical VK and
    identical proof bytes under both transcripts and both multi-open arguments."""
    t = KAT["proof_k5"]
    ctx = zkw.Context(0)
    try:
        k = t["shape"][0]
        ctx.srs_setup(k, oracle.fr_to_mont([TAU])[0])
        shape = zkw.CircuitShape.from_config(k, 1, 1, 1)
        fixed = [oracle.fr_to_mont(ints(col)) for col in t["fixed"]]
        mapping = [np.array(m, dtype=np.uint32) for m in t["mapping"]]
        pk = zkw.keygen(ctx, shape, fixed, mapping)
        fx, pm, dg = pk.vk()
        assert [list(oracle.g1_affine_to_ints(p)) for p in fx] == [ints(p) for p in t["vk_fixed"]]
        assert [list(oracle.g1_affine_to_ints(p)) for p in pm] == [ints(p) for p in t["vk_perm"]]
        assert oracle.fr_from_mont(dg.reshape(1, 4))[0] == int(t["vk_digest"], 16)
        advice = [oracle.fr_to_mont(ints(col)) for col in t["advice"]]
        for key, hx in t["proofs"].items():
            kind, mo = key.split("/")
            tr = zkw.TRANSCRIPT_EVM if kind == "evm" else zkw.TRANSCRIPT_BLAKE2B
            got = zkw.create_proof(ctx, pk, advice, seed=t["blinding_seed"], transcript=tr, shplonk=(mo == "shplonk"))
            assert got.hex() == hx, key
        pk.close()
    finally:
        ctx.close()
