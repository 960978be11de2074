"""Short targets for `ncu --set full` (development aid, see profiles/):
  msm   : device SRS setup (2^19) + three uniform-scalar MSMs over the resident window tables
  ntt   : two iNTT(2^21) and two coset extensions 2^19 -> 2^21 (no keygen: every ntt_pass launch is one of these)
  quot  : one quotient evaluation at the k = 19 shape on random cosets"""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
zkw = importlib.import_module("webauthn-halo2_b200")
mode = sys.argv[1] if len(sys.argv) > 1 else "msm"
ctx = zkw.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=0); torch.cuda.set_stream(stream)
k = 19; n = 1 << k
def rnd(m):
    t = torch.randint(0, 1 << 62, (m, 4), dtype=torch.int64, device="cuda"); t[:, 3] &= (1 << 60) - 1
    return t
if mode == "msm":
    tau = zkw.prover.fr_to_mont(ctx, np.array([[0x1234567890ABCDEF, 0x1234567890ABCDEF, 0x1234567890ABCDEF, 0]], dtype=np.uint64))[0]
    ctx.srs_setup(k, tau)
    s = rnd(n)
    for _ in range(3):
        ctx.msm_dev(s, n, zkw.BASES_G)
elif mode == "ntt":
    a = rnd(4 * n); c = rnd(n); e = torch.empty((4 * n, 4), dtype=torch.int64, device="cuda")
    for _ in range(2):
        ctx.lagrange_to_coeff_dev(a, k + 2)
    for _ in range(2):
        ctx.coeff_to_extended_dev(c, k, k + 2, e)
elif mode == "quot":
    shape = zkw.CircuitShape.from_config(k, 1, 1, 1)
    en = 4 * n
    dv = lambda: rnd(en)
    cols = {"advice": [dv()], "constants": [dv()], "table": dv(), "q_enable": [dv()], "q_lookup": dv(), "sigma": [dv(), dv()],
            "perm_z": [dv()], "lookup_z": [dv()], "lookup_a": [dv()], "lookup_s": [dv()], "l0": dv(), "l_last": dv(), "l_active": dv()}
    ch = {nme: [3, 5, 7, 11] for nme in ("y", "beta", "gamma", "theta")}
    ch = {k_: np.array(v, dtype=np.uint64) for k_, v in ch.items()}
    h = torch.empty((en, 4), dtype=torch.int64, device="cuda")
    for _ in range(2):
        ctx.quotient_dev(shape, cols, ch, h)
ctx.sync()
print("done", mode)
