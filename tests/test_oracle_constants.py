"""Pins the oracle's field / domain constants to the values the reference embeds in its generated
verifier (proving-server/P256Verifier.yul).  The constants are committed here as a golden fixture
(tests/golden/yul_constants.json, extracted by tools/extract_yul_constants.py) because
/root/reference does not travel; when the reference IS mounted the fixture is re-checked against it."""
import json
import os

import pytest

from oracle import pyref as pr

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "yul_constants.json")))


def test_moduli_match_yul():
    assert pr.P == int(GOLD["f_p"], 16)      # yul:17
    assert pr.R == int(GOLD["f_q"], 16)      # yul:18


def test_domain_constants_k17_match_yul():
    dom = pr.EvaluationDomain(4, 17)
    assert dom.extended_k == 19
    assert pow(1 << 17, -1, pr.R) == int(GOLD["n_inv_k17"])                      # yul:307
    # -omega^-j for the Lagrange set {omega^0, omega^-1 .. omega^-7} (yul:308-323), stored as (j, omega^-j, -omega^-j)
    for j, w, negw in GOLD["omega_powers_k17"]:
        assert pow(dom.omega, -j, pr.R) == int(w)
        assert (-pow(dom.omega, -j, pr.R)) % pr.R == int(negw)


def test_delta_powers_match_yul():
    # delta = 7^(2^28): permutation argument's coset shifts delta^0..delta^5 (yul:465,483,487,505,509)
    for i, d in enumerate(GOLD["delta_powers"], start=0):
        assert pow(pr.FR_DELTA, i, pr.R) == int(d)


def test_zeta_is_a_primitive_cube_root_and_matches_halo2curves():
    assert pow(pr.FR_ZETA, 3, pr.R) == 1 and pr.FR_ZETA != 1
    assert pr.FR_ZETA == 0xB3C4D79D41A917585BFC41088D8DAAA78B17EA66B99C90DD


def test_root_of_unity_order():
    assert pow(pr.FR_ROOT_OF_UNITY, 1 << 28, pr.R) == 1
    assert pow(pr.FR_ROOT_OF_UNITY, 1 << 27, pr.R) == pr.R - 1


@pytest.mark.skipif(not os.path.exists("/root/reference/proving-server/P256Verifier.yul"), reason="reference not mounted")
def test_fixture_matches_mounted_reference():
    import subprocess, sys
    out = subprocess.run([sys.executable, os.path.join(HERE, "..", "tools", "extract_yul_constants.py"), "--stdout"],
                         check=True, capture_output=True, text=True).stdout
    assert json.loads(out) == GOLD
