"""Times iNTT(2^19), iNTT(2^21) and the coset extension 2^19 -> 2^21 for several stage splits
(ZKW_NTT_PLAN_<log_n>, see ntt.cu).  Development aid."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
zkw = importlib.import_module("webauthn-halo2_b200")
ctx = zkw.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=0); torch.cuda.set_stream(stream)
def timed(fn, reps=20):
    fn(); stream.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps): fn()
    b.record(stream); b.synchronize()
    return a.elapsed_time(b) / reps
def rnd(n):
    t = torch.randint(0, 1 << 62, (n, 4), dtype=torch.int64, device="cuda"); t[:, 3] &= (1 << 60) - 1
    return t
a19, a21 = rnd(1 << 19), rnd(1 << 21)
e = torch.empty((1 << 21, 4), dtype=torch.int64, device="cuda")
plans19 = ["7,6,6", "6,7,6", "6,6,7", "9,10", "10,9", "9,5,5", "4,6,9", "6,4,9", "3,8,8", "8,8,3", "5,5,9", "9,9,1", "6,6,6,1"]
plans21 = ["7,7,7", "9,6,6", "6,9,6", "6,6,9", "3,9,9", "9,9,3", "9,3,9", "8,7,6", "6,7,8", "9,7,5", "5,7,9", "6,6,6,3", "10,10,1", "10,6,5"]
for p in plans19:
    os.environ["ZKW_NTT_PLAN_19"] = p
    print("2^19 plan %-8s intt %.4f ms" % (p, timed(lambda: ctx.lagrange_to_coeff_dev(a19, 19))), flush=True)
os.environ.pop("ZKW_NTT_PLAN_19")
for p in plans21:
    os.environ["ZKW_NTT_PLAN_21"] = p
    print("2^21 plan %-8s intt %.4f ms   coset 19->21 %.4f ms" % (p, timed(lambda: ctx.lagrange_to_coeff_dev(a21, 21)), timed(lambda: ctx.coeff_to_extended_dev(a19, 19, 21, e))), flush=True)
