"""Per-launch timeline of one proof (k = 19 by default; second argument = degree) (CUDA events around every kernel, all streams), written as CSV:
name,stream,start_ms,duration_ms.  Development aid: shows what overlaps what and where the device idles."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline.csv"
degree = int(sys.argv[2]) if len(sys.argv) > 2 else 19
if os.path.exists(out):
    os.remove(out)
zkw = importlib.import_module("webauthn-halo2_b200")
st = zkw.ProverState(zkw.CircuitParams.for_degree(degree), 0)
ctx = st.ctx
_a = zkw.synthetic_assertion(1)
cols = st.circuit.synthesize(*[_a[32 * j: 32 * j + 32] for j in range(5)])
dev = [torch.from_numpy(c.view(np.int64)).cuda() for c in cols]
rows = [c.shape[0] for c in cols]
def prove(seed):
    return zkw.create_proof(ctx, st.pk, dev, seed=seed, transcript=zkw.TRANSCRIPT_EVM, canonical=True, device_rows=rows)
prove(1); prove(2)
ctx.sync()
os.environ["ZKW_TIMELINE"] = out
ctx.profile_enable(True)
ctx.profile_reset()
prove(3)
tot = ctx.profile_all()
ctx.profile_enable(False)
rowsz = [l.strip().split(",") for l in open(out)]
t_end = max(float(r[2]) + float(r[3]) for r in rowsz)
print(f"{len(rowsz)} launches, span {t_end:.2f} ms (profiling events add overhead)")
