"""MSM time vs window width c for large N (resident SRS with window tables).  Development aid for pick_window_bits."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
zkw = importlib.import_module("webauthn-halo2_b200")
tau = np.array([0x1234567890ABCDEF, 0x0FEDCBA987654321, 0x1111111111111111, 0x0222222222222222], dtype=np.uint64)
for k in (21, 22, 23, 24):
    n = 1 << k
    res = []
    for c in range(16, 24):
        ctx = zkw.Context(0)
        ctx.msm_config(c, True)
        try:
            ctx.srs_setup(k, tau)
            stream = torch.cuda.ExternalStream(ctx.stream, device=0); torch.cuda.set_stream(stream)
            s = torch.randint(0, 1 << 62, (n, 4), dtype=torch.int64, device="cuda"); s[:, 3] &= (1 << 60) - 1
            ctx.msm_dev(s, n, zkw.BASES_G); stream.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(3): ctx.msm_dev(s, n, zkw.BASES_G)
            b.record(stream); b.synchronize()
            res.append((c, a.elapsed_time(b) / 3))
        except Exception as ex:
            res.append((c, float("nan")))
        ctx.close(); torch.cuda.empty_cache()
    print("k=%d " % k + "  ".join("c=%d:%.3f" % r for r in res), flush=True)
