import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
zkw = importlib.import_module("webauthn-halo2_b200")
st = zkw.ProverState(zkw.CircuitParams.for_degree(19), 0)
for i in range(3): st.prove(zkw.synthetic_assertion(i), zkw.TRANSCRIPT_EVM, seed=i)
stg = st._staging
ts = []
for i in range(10):
    t0 = time.perf_counter(); zkw.native.synth_witness(st.shape, st.params.lookup_bits, b"a%d" % i, out=stg); ts.append(time.perf_counter() - t0)
print("synth_witness ms:", [round(x * 1e3, 3) for x in ts])
ts = []
for i in range(10):
    t0 = time.perf_counter(); p = zkw.create_proof(st.ctx, st.pk, stg, i, zkw.TRANSCRIPT_EVM, u64=True); ts.append(time.perf_counter() - t0)
print("create_proof(host u64 cols) ms:", [round(x * 1e3, 2) for x in ts])
ts = []
for i in range(10):
    t0 = time.perf_counter(); p = st.prove(zkw.synthetic_assertion(10 + i), zkw.TRANSCRIPT_EVM, seed=i); ts.append(time.perf_counter() - t0)
print("prove ms:", [round(x * 1e3, 2) for x in ts])
print(os.cpu_count())
