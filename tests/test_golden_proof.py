"""The reference's golden EVM proof (contracts/test/P256Account.t.sol:120, 2720 bytes; the reference
asserts its acceptance at :89-101) against (1) the reference's own generated verifier executed by the
oracle's Yul interpreter and (2) the oracle's Python restatement of halo2's verify_proof + GWC.

(1) pins the oracle's keccak256, BN254 G1/G2 arithmetic and pairing; (2) pins the restated verification
algorithm — proof layout, transcript framing, constraint list and y-ordering, rotation sets and v/u
ordering, final pairing equation — to the one the reference ships.  The device prover's proofs are
then checked with that same verifier (tests/test_gpu_prover.py)."""
import json
import os

import pytest

from oracle import halo2_ref as h, pairing as pg, pyref as pr

HERE = os.path.dirname(os.path.abspath(__file__))
YUL = "/root/reference/proving-server/P256Verifier.yul"
PROOF = bytes.fromhex(open(os.path.join(HERE, "golden", "golden_proof_k17_evm.hex")).read().strip())
VKJ = json.load(open(os.path.join(HERE, "golden", "vk_k17_evm.json")))


def _vk():
    w = [int(x, 16) for x in VKJ["vk_points_xy"]]
    pts = [(w[i], w[i + 1]) for i in range(0, len(w), 2)]
    assert pts[0] == pr.G1_GEN and len(pts) == 13 and all(pr.g1_is_on_curve(p) for p in pts)

    def g2dec(ws):
        a = [int(x, 16) for x in ws]
        return ((a[1], a[0]), (a[3], a[2]))  # EIP-197 order: imaginary part first

    g2 = g2dec(VKJ["g2"])
    n = g2dec(VKJ["neg_s_g2"])
    s_g2 = (n[0], ((-n[1][0]) % pr.P, (-n[1][1]) % pr.P))
    assert g2 == pg.G2_GEN and pg.g2_is_on_curve(s_g2)
    shape = h.Shape(17, 4, 1, 1)  # halo2-circuits/src/configs/ecdsa_circuit.config:1
    return h.VerifyingKey(shape, pts[1:7], pts[7:13], int(VKJ["vk_digest"]), pts[0]), (g2, s_g2)


def test_golden_proof_shape():
    assert len(PROOF) == 2720
    shape = h.Shape(17, 4, 1, 1)
    points = shape.num_advice_cols + 2 * shape.num_lookups + shape.num_perm_sets + shape.num_lookups + 1 + shape.quotient_pieces + 6
    evals = len(shape.advice_queries()) + len(shape.fixed_queries()) + 1 + len(shape.perm_columns()) + (3 * shape.num_perm_sets - 1) + 5 * shape.num_lookups
    assert (points, evals) == (21, 43) and 64 * points + 32 * evals == 2720


def test_python_verifier_accepts_reference_golden_proof():
    vk, g2_pair = _vk()
    assert h.verify_proof(vk, PROOF, "evm", g2_pair=g2_pair)


@pytest.mark.parametrize("pos", [5, 700, 1400, 2000, 2719])
def test_python_verifier_rejects_tampered_golden_proof(pos):
    vk, g2_pair = _vk()
    bad = bytearray(PROOF)
    bad[pos] ^= 1
    assert not h.verify_proof(vk, bytes(bad), "evm", g2_pair=g2_pair)


def test_python_verifier_rejects_truncated_and_empty():
    vk, g2_pair = _vk()
    assert not h.verify_proof(vk, PROOF[:-64], "evm", g2_pair=g2_pair)
    assert not h.verify_proof(vk, b"", "evm", g2_pair=g2_pair)   # the reference's empty-signature case (P256Account.t.sol:106-118)


@pytest.mark.skipif(not os.path.exists(YUL), reason="reference not mounted (the Yul is not copied into the repo)")
def test_reference_yul_verifier_accepts_golden_proof_under_oracle_evm():
    from oracle import yul_evm
    src = open(YUL).read()
    ok, m = yul_evm.run_verifier(src, PROOF)
    assert ok
    assert m.precompile_calls[8] == 1 and m.keccak_calls == 7
    bad = bytearray(PROOF)
    bad[100] ^= 1
    assert not yul_evm.run_verifier(src, bytes(bad))[0]
    assert not yul_evm.run_verifier(src, b"")[0]


@pytest.mark.skipif(not os.path.exists(YUL), reason="reference not mounted")
def test_python_verifier_and_yul_agree_on_challenges():
    """theta, beta, gamma, y, x, v, u as the Yul derives them (memory 0x180, 0x260, 0x2c0, 0x460, 0x580,
    0xb40, 0xd20) equal the restated EvmTranscript's."""
    from oracle import yul_evm
    _, m = yul_evm.run_verifier(open(YUL).read(), PROOF)
    tr = h.EvmTranscript(PROOF)
    tr.common_scalar(int(VKJ["vk_digest"]))
    got = []
    for npts, nsq in ((5, 1), (2, 2), (5, 1), (3, 1)):
        for _ in range(npts):
            tr.read_point()
        got += [tr.squeeze() for _ in range(nsq)]
    for _ in range(43):
        tr.read_scalar()
    got.append(tr.squeeze())
    for _ in range(6):
        tr.read_point()
    got.append(tr.squeeze())
    assert got == [m.mload(a) for a in (0x180, 0x260, 0x2C0, 0x460, 0x580, 0xB40, 0xD20)]
