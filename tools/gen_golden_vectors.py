"""Generates tests/golden/kat_small.json: small known-answer vectors for the hot-path functions and for
whole proofs, produced by the oracle (C restatement + Python prover) from seeded inputs.  The reference
itself holds no such vectors and cannot be run here, so these pin the ORACLE's current
behaviour — which is itself pinned to the reference through the golden proof (tests/test_golden_proof.py) —
and let both the CPU suite (oracle == fixture) and the GPU suite (device == fixture) detect regressions
without trusting a live re-computation.  Run in the build container: python tools/gen_golden_vectors.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import cpu, halo2_ref as h, pyref as pr, synth_circuit as sc  # noqa: E402

TAU = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF


def hexs(vals):
    return ["%064x" % v for v in vals]


def main():
    out = {"note": "oracle-generated known answers; canonical integers as 64-digit hex", "tau": "%x" % TAU}
    # NTT, n = 16: input = fr_random(16, seed 11), omega = 16th root
    a = cpu.fr_random(16, 11)
    w = pow(pr.FR_ROOT_OF_UNITY, 1 << (28 - 4), pr.R)
    out["ntt16"] = {"seed": 11, "omega": "%064x" % w, "input": hexs(cpu.fr_from_mont(a)),
                    "output": hexs(cpu.fr_from_mont(cpu.best_fft(a, cpu.fr_to_mont([w])[0])))}
    # coset extension k = 3 -> 5 (degree-4 system) and back
    d = cpu.Domain.new(4, 3)
    c = cpu.fr_random(8, 12)
    out["coset_ext_k3"] = {"seed": 12, "coeffs": hexs(cpu.fr_from_mont(c)), "extended": hexs(cpu.fr_from_mont(d.coeff_to_extended(c)))}
    # MSM, n = 8: scalars fr_random(8, 13), bases = fr_random(8, 14) * G
    s = cpu.fr_random(8, 13)
    b = cpu.g1_fixed_base_mul(cpu.fr_random(8, 14))
    res = cpu.g1_affine_to_ints(cpu.g1_to_affine(cpu.best_multiexp(s, b))[0])
    out["msm8"] = {"scalar_seed": 13, "base_scalar_seed": 14, "result_xy": hexs(res)}
    # SRS: first 4 powers of tau
    out["srs_g4"] = [hexs(cpu.g1_affine_to_ints(p)) for p in cpu.srs_powers(4, cpu.fr_to_mont([TAU])[0])]
    # whole proofs from the Python oracle prover on the oracle's synthetic circuit (seed 5), k = 5, selector mode
    shape = h.Shape(5, 1, 0, 1)
    fixed, mapping, advice = sc.build(shape, seed=5)
    n = shape.n
    dom = shape.domain()
    g = cpu.srs_powers(n, cpu.fr_to_mont([TAU])[0])
    cst = (pow(TAU, n, pr.R) - 1) * pow(n, -1, pr.R) % pr.R
    ls = [cst * pow(dom.omega, i, pr.R) % pr.R * pow((TAU - pow(dom.omega, i, pr.R)) % pr.R, -1, pr.R) % pr.R for i in range(n)]
    gl = cpu.g1_fixed_base_mul(cpu.fr_to_mont(ls))
    pk = h.keygen(shape, gl, fixed, h.sigma_from_cycles(shape, mapping))
    proofs = {}
    for kind in ("evm", "blake2b"):
        for mo in ("gwc", "shplonk"):
            p = h.create_proof(pk, g, gl, advice, seed=2024, kind=kind, multiopen=mo)
            assert h.verify_proof(pk.vk, p, kind, tau=TAU, multiopen=mo)
            proofs[f"{kind}/{mo}"] = p.hex()
    out["proof_k5"] = {
        "shape": [5, 1, 0, 1], "circuit_seed": 5, "blinding_seed": 2024,
        "fixed": [hexs(col) for col in fixed], "mapping": mapping, "advice": [hexs(col) for col in advice],
        "vk_digest": "%064x" % pk.vk.digest,
        "vk_fixed": [hexs(p) for p in pk.vk.fixed_commitments], "vk_perm": [hexs(p) for p in pk.vk.perm_commitments],
        "proofs": proofs,
    }
    path = os.path.join(ROOT, "tests", "golden", "kat_small.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
