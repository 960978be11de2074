"""Uniform-scalar MSM over resident window tables for several (log n, window width c) pairs, and the k = 17 proof per c
(development aid; the default is chosen in msm.cu pick_window_bits):  python tools/window_sweep_sizes.py 17:16,17 18:16,17 ..."""
import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
zkw = importlib.import_module("webauthn-halo2_b200")
for spec in sys.argv[1:] or ["17:16,17", "18:16,17", "20:17,18", "21:17,20"]:
    k, cs = spec.split(":")
    k = int(k); n = 1 << k
    for c in [int(x) for x in cs.split(",")]:
        ctx = zkw.Context(0)
        ctx.msm_config(c, True)
        tau = zkw.prover.fr_to_mont(ctx, np.array([[0x1234567890ABCDEF, 0x1234567890ABCDEF, 0x1234567890ABCDEF, 0]], dtype=np.uint64))[0]
        ctx.srs_setup(k, tau)
        stream = torch.cuda.ExternalStream(ctx.stream, device=0); torch.cuda.set_stream(stream)
        s = torch.randint(0, 1 << 62, (n, 4), dtype=torch.int64, device="cuda"); s[:, 3] &= (1 << 60) - 1
        def timed(fn, reps=10):
            fn(); stream.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(reps): fn()
            b.record(stream); b.synchronize()
            return a.elapsed_time(b) / reps
        print("k=%d c=%d  msm %.3f ms" % (k, c, timed(lambda: ctx.msm_dev(s, n, zkw.BASES_G))), flush=True)
        ctx.close(); del s
        torch.cuda.empty_cache()
