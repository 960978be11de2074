"""Quick device timings (CUDA events on the ctx stream) for the three hot-path kernels at the k=19
shape.  Development aid; bench.py is the contract."""
import importlib
import sys
import os
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
zkw = importlib.import_module("webauthn-halo2_b200")
from oracle import cpu as oracle


def timed(stream, fn, reps=5):
    fn()
    stream.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sum(ts) / len(ts)


def main():
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 19
    ek = k + 2
    n, en = 1 << k, 1 << ek
    torch.cuda.init()
    ctx = zkw.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream)
    t0 = time.time()
    g = oracle.g1_fixed_base_mul(oracle.fr_random(n, 1))
    print(f"srs gen (cpu) {time.time()-t0:.2f}s")
    t0 = time.time()
    ctx.srs_load(g, None)
    print(f"srs_load + window tables {time.time()-t0:.3f}s")
    s = torch.from_numpy(oracle.fr_random(n, 2).view(np.int64)).cuda()
    for label, fn in (
        ("msm 2^%d (tables)" % k, lambda: ctx.msm_dev(s, n, zkw.BASES_G)),
    ):
        best, avg = timed(stream, fn)
        print(f"{label}: best {best:.3f} ms avg {avg:.3f} ms  -> {96*n/best/1e6:.1f} GB/s algorithmic")
    a = torch.from_numpy(oracle.fr_random(n, 3).view(np.int64)).cuda()
    e = torch.empty((en, 4), dtype=torch.int64, device="cuda")
    best, avg = timed(stream, lambda: ctx.lagrange_to_coeff_dev(a, k))
    print(f"iNTT 2^{k}: best {best:.3f} ms avg {avg:.3f}  -> {64*n/best/1e6:.1f} GB/s algorithmic")
    best, avg = timed(stream, lambda: ctx.coeff_to_extended_dev(a, k, ek, e))
    print(f"coeff_to_extended 2^{k}->2^{ek}: best {best:.3f} ms avg {avg:.3f} -> {(32*n+32*en)/best/1e6:.1f} GB/s")
    best, avg = timed(stream, lambda: ctx.extended_to_coeff_dev(e, ek))
    print(f"extended_to_coeff 2^{ek}: best {best:.3f} ms avg {avg:.3f} -> {64*en/best/1e6:.1f} GB/s")
    # quotient at the k=19 shape with random cosets
    shape = zkw.CircuitShape.from_config(k, 1, 1, 1)
    def dv():
        return torch.randint(0, 1 << 60, (en, 4), dtype=torch.int64, device="cuda")
    cols = {"advice": [dv()], "constants": [dv()], "table": dv(), "q_enable": [dv()], "q_lookup": dv(),
            "sigma": [dv(), dv()], "perm_z": [dv()], "lookup_z": [dv()], "lookup_a": [dv()], "lookup_s": [dv()],
            "l0": dv(), "l_last": dv(), "l_active": dv()}
    ch = {nme: oracle.fr_random(1, 9)[0] for nme in ("y", "beta", "gamma", "theta")}
    h = torch.empty((en, 4), dtype=torch.int64, device="cuda")
    best, avg = timed(stream, lambda: ctx.quotient_dev(shape, cols, ch, h))
    print(f"quotient 2^{ek} rows: best {best:.3f} ms avg {avg:.3f} -> {32*en*15/best/1e6:.1f} GB/s algorithmic")
    print("launches", ctx.launch_count)


if __name__ == "__main__":
    main()
