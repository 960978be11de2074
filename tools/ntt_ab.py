"""Device-resident NTT timings (development aid for A/B runs through ZKW_B200_LIB): coset extension 2^k -> 2^(k+2) and the
inverse transform at 2^k, CUDA events on the context's stream, median of 20."""
import importlib, os, sys
import torch
sys.path.insert(0, os.getcwd())
zkw = importlib.import_module("webauthn-halo2_b200")
ctx = zkw.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=0); torch.cuda.set_stream(stream)
def timed(fn, reps=20):
    for _ in range(3): fn()
    stream.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream); b.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[reps // 2]
out = []
for k in [int(x) for x in sys.argv[1:]] or [17, 19]:
    src = torch.randint(0, 1 << 62, (1 << k, 4), dtype=torch.int64, device="cuda"); src[:, 3] &= (1 << 59) - 1
    ext = torch.empty((1 << (k + 2), 4), dtype=torch.int64, device="cuda")
    out.append("k=%d coset %.4f ms  intt %.4f ms" % (k, timed(lambda: ctx.coeff_to_extended_dev(src, k, k + 2, ext)),
                                                     timed(lambda: ctx.lagrange_to_coeff_dev(src, k))))
print(os.environ.get("ZKW_B200_LIB", "default")[-14:], " | ".join(out))
