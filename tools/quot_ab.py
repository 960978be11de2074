import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
zkw = importlib.import_module("webauthn-halo2_b200")
ctx = zkw.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=0); torch.cuda.set_stream(stream)
def rnd(m):
    t = torch.randint(0, 1 << 62, (m, 4), dtype=torch.int64, device="cuda"); t[:, 3] &= (1 << 60) - 1
    return t
out=[]
for (k,A,L) in ((19,1,1),(17,4,1)):
    shape = zkw.CircuitShape.from_config(k, A, L, 1)
    en = 1 << shape.ext_k
    nl = shape.num_lookup_advice; ncols = 1 + A + nl; nsets = shape.perm_sets; nlk = shape.lookups
    dv = lambda: rnd(en)
    cols = {"advice": [dv() for _ in range(A+nl)], "constants": [dv()], "table": dv(), "q_enable": [dv() for _ in range(A)], "q_lookup": dv() if nl==0 else None,
            "sigma": [dv() for _ in range(ncols)], "perm_z": [dv() for _ in range(nsets)], "lookup_z": [dv() for _ in range(nlk)], "lookup_a": [dv() for _ in range(nlk)],
            "lookup_s": [dv() for _ in range(nlk)], "l0": dv(), "l_last": dv(), "l_active": dv()}
    ch = {n_: np.array([3+i, 5, 7, 11], dtype=np.uint64) for i, n_ in enumerate(("y", "beta", "gamma", "theta"))}
    h = torch.empty((en, 4), dtype=torch.int64, device="cuda")
    for _ in range(3): ctx.quotient_dev(shape, cols, ch, h)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream.synchronize(); a.record(stream)
    for _ in range(10): ctx.quotient_dev(shape, cols, ch, h)
    b.record(stream); b.synchronize()
    out.append("k=%d quotient %.4f ms" % (k, a.elapsed_time(b) / 10))
print(os.environ.get("ZKW_B200_LIB", "default")[-12:], " | ".join(out))
