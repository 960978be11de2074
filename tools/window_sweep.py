"""Times a uniform-scalar MSM (resident SRS + window tables, 2^19) and a whole k=19 proof for several MSM
window widths c (development aid; the default is chosen in msm.cu pick_window_bits)."""
import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
zkw = importlib.import_module("webauthn-halo2_b200")
n = 1 << 19
for c in [int(x) for x in (sys.argv[1:] or ["16", "17", "18", "19"])]:
    ctx = zkw.Context(0)
    ctx.msm_config(c, True)
    st = zkw.ProverState(zkw.CircuitParams.for_degree(19), 0, ctx)
    stream = torch.cuda.ExternalStream(ctx.stream, device=0); torch.cuda.set_stream(stream)
    s = torch.randint(0, 1 << 62, (n, 4), dtype=torch.int64, device="cuda"); s[:, 3] &= (1 << 60) - 1
    def timed(fn, reps=10):
        fn(); stream.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps): fn()
        b.record(stream); b.synchronize()
        return a.elapsed_time(b) / reps
    msm = timed(lambda: ctx.msm_dev(s, n, zkw.BASES_G))
    cols = st.circuit.synthesize(b"a")
    dev = [torch.from_numpy(zkw.circuit.to_limbs(cc).view(np.int64)).cuda() for cc in cols]
    rows = [cc.shape[0] for cc in cols]
    i = [0]
    def prove():
        i[0] += 1
        return zkw.create_proof(ctx, st.pk, dev, seed=i[0], transcript=zkw.TRANSCRIPT_EVM, canonical=True, device_rows=rows)
    prove(); prove()
    t0 = time.perf_counter()
    for _ in range(6): prove()
    pt = (time.perf_counter() - t0) / 6 * 1e3
    print("c=%d  msm %.3f ms  proof %.2f ms" % (c, msm, pt), flush=True)
    st.close(); ctx.close()
    torch.cuda.empty_cache()
