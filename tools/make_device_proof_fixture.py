"""Promotes the k = 17 EVM/GWC device proof that `pytest -m gpu` leaves in gpurun_out/device_proof_k17_evm.json
(tests/test_gpu_prover.py::test_k17_evm_device_proof_accepted_with_real_pairing: real ECDSA circuit, signed
assertion 17, blinding seed 5, development tau) to the committed fixture tests/golden/device_proof_k17_evm.json,
after checking here - where /root/reference is mounted - that the reference's own Yul verifier
(proving-server/P256Verifier.yul) with this key's constants swapped in accepts it.  The GPU test then asserts the
device still produces exactly these bytes, and tests/test_device_proof_fixture.py re-runs the Yul check on CPU."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyref as pr, yul_evm, yul_patch  # noqa: E402

SRC = os.path.join(ROOT, "gpurun_out", "device_proof_k17_evm.json")
DST = os.path.join(ROOT, "tests", "golden", "device_proof_k17_evm.json")
YUL = "/root/reference/proving-server/P256Verifier.yul"


def main():
    d = json.load(open(SRC))
    proof = bytes.fromhex(d["proof"])
    fx = [(int(a, 16), int(b, 16)) for a, b in d["fixed"]]
    pm = [(int(a, 16), int(b, 16)) for a, b in d["perm"]]
    src = yul_patch.patch_verifier(open(YUL).read(), int(d["digest"]), pr.G1_GEN, fx, pm, int(d["tau"], 16))
    ok, m = yul_evm.run_verifier(src, proof)
    assert ok and m.precompile_calls[8] == 1, "the reference verifier rejects the device proof"
    d["source"] = ("device prover (libzkw_b200.so) on a B200: generate_proof_evm(signed_assertion(17), degree 17, seed 5); accepted by "
                   "proving-server/P256Verifier.yul with digest / 12 VK points / -s*G2 swapped (yul:34,880-980,1131-1134)")
    with open(DST, "w") as f:
        json.dump(d, f, indent=1)
    print("fixture written:", DST, len(proof), "bytes; pairing precompile calls:", m.precompile_calls[8])


if __name__ == "__main__":
    main()
