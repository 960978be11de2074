"""Resident-witness proof time at the given degrees (development aid for A/B runs of environment knobs or of a second
build through ZKW_B200_LIB):  python tools/proof_ab.py 17 19  -> median / min ms of 15 proofs per degree, CUDA events."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
zkw = importlib.import_module("webauthn-halo2_b200")
degrees = [int(a) for a in sys.argv[1:]] or [17, 19]
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("ZKW_"))
for degree in degrees:
    st = zkw.ProverState(zkw.CircuitParams.for_degree(degree), 0)
    ctx = st.ctx
    a = zkw.synthetic_assertion(1)
    cols = st.circuit.synthesize(*[a[32 * j: 32 * j + 32] for j in range(5)])
    dev = [torch.from_numpy(c.view(np.int64)).cuda() for c in cols]
    rows = [c.shape[0] for c in cols]
    def prove(seed):
        return zkw.create_proof(ctx, st.pk, dev, seed=seed, transcript=zkw.TRANSCRIPT_EVM, canonical=True, device_rows=rows)
    for i in range(4):
        prove(i)
    ts = []
    for i in range(15):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); prove(100 + i); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"k={degree} [{tag}] median {ts[len(ts) // 2]:.3f} ms  min {ts[0]:.3f} ms", flush=True)
    del st, ctx, dev
