"""TEST INFRASTRUCTURE ONLY.  Swaps the verifying-key constants embedded in the reference's generated Yul
verifier (proving-server/P256Verifier.yul: transcript digest :34, G1 generator + 12 fixed / permutation
commitments :880-980, -s*G2 :1131-1134) for another key's, leaving every instruction of the verification
algorithm untouched.  With the constants of THIS repo's k = 17 proving key the reference's own verifier then
judges the device prover's proofs.  Nothing here is shipped; the Yul text itself is read from
/root/reference at test time and never copied into the repo."""
from __future__ import annotations

import re

from . import pairing as pg
from .pyref import P

_CONST = re.compile(r"(mstore\(0x[0-9a-f]+, )(0x[0-9a-f]{64})(\))")
_DIGEST = re.compile(r"(mstore\(0x0, )(\d+)(\))")


def neg_s_g2_words(tau: int):
    """-tau*G2 as the four EIP-197 words (x_im, x_re, y_im, y_re) the verifier stores (yul:1131-1134)."""
    s_g2 = pg.g2_mul(pg.G2_GEN, tau)
    (x_re, x_im), (y_re, y_im) = s_g2
    return [x_im, x_re, (-y_im) % P, (-y_re) % P], s_g2


def patch_verifier(yul_src: str, digest: int, g0, fixed_commitments, perm_commitments, tau: int) -> str:
    """Returns the Yul text with digest, the 13 G1 constants (generator, fixed.., permutation..) and -s*G2
    replaced, in the order the text holds them.  Raises if the text does not have the expected 34 constants."""
    consts = _CONST.findall(yul_src)
    pts = [g0] + list(fixed_commitments) + list(perm_commitments)
    if len(consts) != 2 * len(pts) + 8:
        raise ValueError(f"expected {2 * len(pts) + 8} embedded constants, found {len(consts)}")
    words = []
    for x, y in pts:
        words += [x, y]
    g2 = pg.G2_GEN
    words += [g2[0][1], g2[0][0], g2[1][1], g2[1][0]]
    neg, _ = neg_s_g2_words(tau)
    words += neg
    it = iter(words)
    out = _CONST.sub(lambda m: f"{m.group(1)}0x{next(it):064x}{m.group(3)}", yul_src)
    out, cnt = _DIGEST.subn(lambda m: f"{m.group(1)}{digest}{m.group(3)}", out, count=1)
    if cnt != 1:
        raise ValueError("transcript digest constant not found")
    return out
