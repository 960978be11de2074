"""Extracts the reference's golden EVM proof (contracts/test/P256Account.t.sol:120, 2720 bytes, accepted
by the reference's verifier in testUserOpE2ESuccess :89-101) and the verifying-key constants embedded in
the generated verifier (proving-server/P256Verifier.yul: vk digest :34, 12 fixed/permutation commitments
:880-980, pairing G2 points :1125-1134) into tests/golden/.  Run in the build container only."""
import json
import os
import re

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def main():
    sol = open(os.path.join(REF, "contracts/test/P256Account.t.sol")).read()
    proof_hex = re.search(r'bytes validSignature =\s*hex"([0-9a-f]+)"', sol).group(1)
    assert len(proof_hex) == 2 * 2720
    with open(os.path.join(OUT, "golden_proof_k17_evm.hex"), "w") as f:
        f.write(proof_hex + "\n")
    yul = open(os.path.join(REF, "proving-server/P256Verifier.yul")).read()
    digest = re.search(r"mstore\(0x0, (\d+)\)", yul).group(1)
    # the 12 VK commitments are written as pairs of 32-byte constants right before the big MSM
    consts = re.findall(r"mstore\(0x[0-9a-f]+, (0x[0-9a-f]{64})\)", yul)
    # last 8 constants: G2 generator (4 words) and -s*G2 (4 words; snark-verifier embeds the NEGATED
    # s*G2 so the check is e(lhs, G2) * e(rhs, -sG2) == 1), each as (x_im, x_re, y_im, y_re)
    g2 = consts[-8:-4]
    sg2 = consts[-4:]
    vk_points = consts[:-8]
    out = {"vk_digest": digest, "vk_points_xy": vk_points, "g2": g2, "neg_s_g2": sg2,
           "source": "proving-server/P256Verifier.yul:34,880-980,1125-1134"}
    with open(os.path.join(OUT, "vk_k17_evm.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("proof bytes", len(proof_hex) // 2, "vk consts", len(vk_points))


if __name__ == "__main__":
    main()
