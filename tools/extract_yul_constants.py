"""Extracts the field / domain constants the reference's generated verifier embeds
(/root/reference/proving-server/P256Verifier.yul) into tests/golden/yul_constants.json.
Run in the build container (the reference is not present on the GPU box)."""
import json
import os
import re
import sys

YUL = "/root/reference/proving-server/P256Verifier.yul"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "yul_constants.json")


def main():
    src = open(YUL).read().splitlines()
    text = "\n".join(src)
    f_p = re.search(r"let f_p := (0x[0-9a-f]+)", text).group(1)   # yul:17
    f_q = re.search(r"let f_q := (0x[0-9a-f]+)", text).group(1)   # yul:18
    # yul:307: mstore(0xfa0, mulmod(mload(0xf80), <n^-1>, f_q))
    n_inv = re.search(r"mstore\(0xfa0, mulmod\(mload\(0xf80\), (\d+), f_q\)\)", text).group(1)
    # yul:308-323: pairs  mulmod(mload(0xfa0), w, f_q)  /  addmod(mload(0x580), -w, f_q)
    ws = re.findall(r"mulmod\(mload\(0xfa0\), (\d+), f_q\)\)\nmstore\(0x[0-9a-f]+, addmod\(mload\(0x580\), (\d+), f_q\)\)", text)
    r = int(f_q, 16)
    assert len(ws) == 8
    # order in the file: omega^-7 (l_last), then omega^-6 .. omega^-1, then omega^0
    powers = []
    for idx, (w, negw) in enumerate(ws):
        j = 7 - idx
        assert (int(w) + int(negw)) % r == 0
        powers.append([j, w, negw])
    # delta powers: constants multiplied by beta (mload(0x260)) in the permutation products
    deltas = re.findall(r"mulmod\((\d+), mload\(0x260\), f_q\)", text)
    out = {"f_p": f_p, "f_q": f_q, "n_inv_k17": n_inv, "omega_powers_k17": powers, "delta_powers": deltas,
           "source": "proving-server/P256Verifier.yul:17-18,307-323,465-509"}
    if "--stdout" in sys.argv:
        print(json.dumps(out))
    else:
        with open(OUT, "w") as f:
            json.dump(out, f, indent=1)
        print("wrote", OUT)


if __name__ == "__main__":
    main()
