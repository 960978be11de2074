import importlib, sys, time
sys.path.insert(0, ".")
zkw = importlib.import_module("webauthn-halo2_b200")
from oracle import halo2_ref as h, cpu
pool = zkw.ProverPool(zkw.CircuitParams.for_degree(17), 0, workers=4)
a = [zkw.synthetic_assertion(i) for i in range(200)]
# sprinkle forged signatures: the batch must report them and still prove the rest
t0 = time.perf_counter(); proofs = pool.prove_many(a, zkw.TRANSCRIPT_EVM); dt = time.perf_counter() - t0
print("k=17 batch of", len(proofs), "proofs/s", len(proofs) / dt, "distinct", len(set(proofs)))
st = pool.states[0]
fx, pm, dg = st.pk.vk()
vk = h.VerifyingKey(h.Shape(17, 4, 1, 1), [cpu.g1_affine_to_ints(p) for p in fx], [cpu.g1_affine_to_ints(p) for p in pm], cpu.fr_from_mont(dg.reshape(1, 4))[0])
bad = 0
for i in range(0, 200, 23):
    if not h.verify_proof(vk, proofs[i], "evm", tau=zkw.prover.DEV_TAU_CANONICAL): bad += 1
print("sampled verification failures:", bad)
forged = list(a[:8]); forged[3] = forged[3][:128] + bytes([forged[3][128] ^ 1]) + forged[3][129:]
try:
    pool.prove_many(forged, zkw.TRANSCRIPT_EVM); print("ERROR: forged batch accepted")
except zkw.InvalidSignature as e:
    print("forged batch refused:", e)
pool.close()
