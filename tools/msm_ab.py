"""A/B helper: times a uniform-scalar MSM (resident SRS, 2^19) and a whole proof with the library named by
ZKW_B200_LIB (development aid)."""
import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
zkw = importlib.import_module("webauthn-halo2_b200")
st = zkw.ProverState(zkw.CircuitParams.for_degree(19), 0)
ctx = st.ctx
stream = torch.cuda.ExternalStream(ctx.stream, device=0); torch.cuda.set_stream(stream)
n = 1 << 19
s = torch.randint(0, 1 << 62, (n, 4), dtype=torch.int64, device="cuda"); s[:, 3] &= (1 << 60) - 1
def timed(fn, reps=10):
    fn(); stream.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps): fn()
    b.record(stream); b.synchronize()
    return a.elapsed_time(b) / reps
msm = timed(lambda: ctx.msm_dev(s, n, zkw.BASES_G))
_a = zkw.synthetic_assertion(1)
cols = st.circuit.synthesize(*[_a[32 * j: 32 * j + 32] for j in range(5)])
dev = [torch.from_numpy(c.view(np.int64)).cuda() for c in cols]
rows = [c.shape[0] for c in cols]
i = [0]
def prove():
    i[0] += 1
    zkw.create_proof(ctx, st.pk, dev, seed=i[0], transcript=zkw.TRANSCRIPT_EVM, canonical=True, device_rows=rows)
prove(); prove()
t0 = time.perf_counter()
for _ in range(8): prove()
pt = (time.perf_counter() - t0) / 8 * 1e3
print(os.environ.get("ZKW_B200_LIB", "default")[-14:], "msm %.3f ms  proof %.2f ms" % (msm, pt))
