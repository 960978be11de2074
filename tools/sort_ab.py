"""A/B helper: the direct entry sort against the binned one (ZKW_MSM_BINNED_SORT, read when a context is created): per-kernel
times of one uniform and one witness-shaped 2^19 MSM, then whole proofs (development aid)."""
import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
zkw = importlib.import_module("webauthn-halo2_b200")
for binned in (0, 1, 0, 1):
    os.environ["ZKW_MSM_BINNED_SORT"] = str(binned)
    st = zkw.ProverState(zkw.CircuitParams.for_degree(19), 0)
    ctx = st.ctx
    stream = torch.cuda.ExternalStream(ctx.stream, device=0); torch.cuda.set_stream(stream)
    n = 1 << 19
    s = torch.randint(0, 1 << 62, (n, 4), dtype=torch.int64, device="cuda"); s[:, 3] &= (1 << 60) - 1
    # a permuted lookup column: sorted 18-bit values, in Montgomery form (v * 2^256 mod r) like every scalar at the boundary
    R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    rng = np.random.default_rng(1)
    vals = [int(v) * (1 << 256) % R_MOD for v in np.sort(rng.integers(0, 1 << 18, n))]
    limbs = np.array([[(v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)] for v in vals], dtype=np.uint64)
    small = torch.from_numpy(limbs.view(np.int64)).cuda()
    def timed(fn, reps=10):
        fn(); stream.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps): fn()
        b.record(stream); b.synchronize()
        return a.elapsed_time(b) / reps
    msm = timed(lambda: ctx.msm_dev(s, n, zkw.BASES_G))
    msm_small = timed(lambda: ctx.msm_dev(small, n, zkw.BASES_G))
    ctx.profile_enable(True); ctx.profile_reset()
    ctx.msm_dev(s, n, zkw.BASES_G); ctx.sync()
    ku = {k: round(v[0] * 1e3) for k, v in ctx.profile_all().items()}
    ctx.profile_reset()
    ctx.msm_dev(small, n, zkw.BASES_G); ctx.sync()
    ks = {k: round(v[0] * 1e3) for k, v in ctx.profile_all().items()}
    ctx.profile_enable(False)
    _a = zkw.synthetic_assertion(1)
    cols = st.circuit.synthesize(*[_a[32 * j: 32 * j + 32] for j in range(5)])
    dev = [torch.from_numpy(c.view(np.int64)).cuda() for c in cols]
    rows = [c.shape[0] for c in cols]
    i = [0]
    def prove():
        i[0] += 1
        return zkw.create_proof(ctx, st.pk, dev, seed=i[0], transcript=zkw.TRANSCRIPT_EVM, canonical=True, device_rows=rows)
    prove(); prove()
    t0 = time.perf_counter()
    for _ in range(10): prove()
    pt = (time.perf_counter() - t0) / 10 * 1e3
    print(f"binned={binned}  uniform msm {msm:.3f} ms  sorted-18-bit msm {msm_small:.3f} ms  proof {pt:.2f} ms")
    print("   uniform us:", ku)
    print("   skewed  us:", ks)
    st.close()
