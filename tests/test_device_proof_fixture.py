"""N1 of the verdict: the reference's OWN verifier judges a proof made by the device prover.

tests/golden/device_proof_k17_evm.json holds a k = 17 EVM/GWC proof of the real P-256 ECDSA circuit produced on a
B200 by libzkw_b200.so (tools/make_device_proof_fixture.py; the GPU suite asserts the device still emits exactly these
bytes) together with this repo's verifying key.  Here, on the CPU:

  * halo2_ref.verify_proof accepts it with the REAL pairing e(left, s*G2) == e(right, G2) — no known-tau shortcut;
  * where /root/reference is mounted, the reference's generated Yul verifier (proving-server/P256Verifier.yul, run by
    oracle/yul_evm.py: own keccak, own BN254 pairing — the interpreter that accepts the reference's golden proof), with
    only the key constants swapped (digest yul:34, 12 commitments yul:880-980, -s*G2 yul:1131-1134), accepts it and
    rejects tampered copies.  The Yul text is read from the reference at test time, never copied into the repo."""
import json
import os

import pytest

from oracle import halo2_ref as h, pairing as pg, pyref as pr

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, "golden", "device_proof_k17_evm.json")
YUL = "/root/reference/proving-server/P256Verifier.yul"

pytestmark = pytest.mark.skipif(not os.path.exists(FIXTURE), reason="fixture not generated yet (tools/make_device_proof_fixture.py)")


def _load():
    d = json.load(open(FIXTURE))
    proof = bytes.fromhex(d["proof"])
    fx = [(int(a, 16), int(b, 16)) for a, b in d["fixed"]]
    pm = [(int(a, 16), int(b, 16)) for a, b in d["perm"]]
    vk = h.VerifyingKey(h.Shape(17, 4, 1, 1), fx, pm, int(d["digest"]))
    return proof, vk, int(d["tau"], 16)


def test_device_proof_has_the_golden_layout():
    proof, vk, _ = _load()
    assert len(proof) == 2720 and len(vk.fixed_commitments) == 6 and len(vk.perm_commitments) == 6
    assert all(pr.g1_is_on_curve(p) for p in vk.fixed_commitments + vk.perm_commitments)


def test_device_proof_verifies_with_the_real_pairing():
    proof, vk, tau = _load()
    g2_pair = (pg.G2_GEN, pg.g2_mul(pg.G2_GEN, tau))
    assert h.verify_proof(vk, proof, "evm", g2_pair=g2_pair)
    bad = bytearray(proof)
    bad[1234] ^= 1
    assert not h.verify_proof(vk, bytes(bad), "evm", g2_pair=g2_pair)
    wrong = (pg.G2_GEN, pg.g2_mul(pg.G2_GEN, tau + 1))        # another SRS: the pairing must fail
    assert not h.verify_proof(vk, proof, "evm", g2_pair=wrong)


@pytest.mark.skipif(not os.path.exists(YUL), reason="reference not mounted (the Yul is not copied into the repo)")
def test_reference_yul_verifier_accepts_the_device_proof():
    from oracle import yul_evm, yul_patch
    proof, vk, tau = _load()
    src = yul_patch.patch_verifier(open(YUL).read(), vk.digest, vk.g0, vk.fixed_commitments, vk.perm_commitments, tau)
    ok, m = yul_evm.run_verifier(src, proof)
    assert ok and m.precompile_calls[8] == 1 and m.keccak_calls == 7
    for pos in (0, 700, 1500, 2719):
        bad = bytearray(proof)
        bad[pos] ^= 1
        assert not yul_evm.run_verifier(src, bytes(bad))[0]
    assert not yul_evm.run_verifier(src, b"")[0]
    # the UNPATCHED reference verifier (the reference's own key) must reject a proof made under another key
    assert not yul_evm.run_verifier(open(YUL).read(), proof)[0]


@pytest.mark.skipif(not os.path.exists(YUL), reason="reference not mounted")
def test_patch_is_the_identity_on_the_reference_key():
    """Swapping in the reference's own constants (tests/golden/vk_k17_evm.json) reproduces the reference's text except for
    -s*G2, whose tau is unknown: the patcher touches nothing but the 34 constants."""
    from oracle import yul_patch
    V = json.load(open(os.path.join(HERE, "golden", "vk_k17_evm.json")))
    w = [int(x, 16) for x in V["vk_points_xy"]]
    pts = [(w[i], w[i + 1]) for i in range(0, len(w), 2)]
    src = open(YUL).read()
    out = yul_patch.patch_verifier(src, int(V["vk_digest"]), pts[0], pts[1:7], pts[7:13], 1)
    a, b = src.splitlines(), out.splitlines()
    diff = [i for i, (x, y) in enumerate(zip(a, b)) if x != y]
    assert len(a) == len(b) and len(diff) <= 4 and all(i >= len(a) - 20 for i in diff)
