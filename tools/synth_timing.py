import importlib, sys, os, time
sys.path.insert(0, os.getcwd())
zkw = importlib.import_module("webauthn-halo2_b200")
k = int(sys.argv[1])
c = zkw.EcdsaCircuit(zkw.CircuitParams.for_degree(k))
a = zkw.synthetic_assertion(1)
args = [a[32*j:32*j+32] for j in range(5)]
cols = c.synthesize(*args)
for i in range(3):
    t0 = time.perf_counter(); c.synthesize(*args, out=cols); print("total %.3f ms" % ((time.perf_counter()-t0)*1e3), file=sys.stderr)
