"""Where a generate_proof_evm call spends its time (development aid):  python tools/e2e_parts.py [degree]"""
import importlib, os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
zkw = importlib.import_module("webauthn-halo2_b200")
st = zkw.ProverState(zkw.CircuitParams.for_degree(int(sys.argv[1]) if len(sys.argv) > 1 else 19), 0)
a = [zkw.synthetic_assertion(i) for i in range(6)]
for i in range(3): st.prove(a[i], zkw.TRANSCRIPT_EVM, seed=i)
# (1) e2e
t0=time.perf_counter()
for i in range(10): st.prove(a[i%6], zkw.TRANSCRIPT_EVM, seed=i)
e2e=(time.perf_counter()-t0)/10*1e3
# (2) device-resident
cols = st.circuit.synthesize(*[a[0][32*j:32*j+32] for j in range(5)])
dev = [torch.from_numpy(c.view(np.int64)).cuda() for c in cols]; rows=[c.shape[0] for c in cols]
for i in range(3): zkw.create_proof(st.ctx, st.pk, dev, seed=i, transcript=zkw.TRANSCRIPT_EVM, canonical=True, device_rows=rows)
t0=time.perf_counter()
for i in range(10): zkw.create_proof(st.ctx, st.pk, dev, seed=i, transcript=zkw.TRANSCRIPT_EVM, canonical=True, device_rows=rows)
res=(time.perf_counter()-t0)/10*1e3
# (3) host columns (pinned staging already synthesised), no synthesis
stg=[st.ctx.host_array(4*r).reshape(r,4) for r in rows]
for s_,c in zip(stg,cols): s_[:]=c
for i in range(3): zkw.create_proof(st.ctx, st.pk, stg, seed=i, transcript=zkw.TRANSCRIPT_EVM, canonical=True)
t0=time.perf_counter()
for i in range(10): zkw.create_proof(st.ctx, st.pk, stg, seed=i, transcript=zkw.TRANSCRIPT_EVM, canonical=True)
h2d=(time.perf_counter()-t0)/10*1e3
# (4) synthesis alone
t0=time.perf_counter()
for i in range(10): st.circuit.synthesize(*[a[i%6][32*j:32*j+32] for j in range(5)], out=stg)
syn=(time.perf_counter()-t0)/10*1e3
print(f"k={sys.argv[1] if len(sys.argv) > 1 else 19} threads={os.environ.get('ZKW_SYNTH_THREADS', 'default')} e2e {e2e:.2f} ms | resident {res:.2f} | pinned-host columns (H2D inside) {h2d:.2f} | synthesis alone {syn:.2f}")
