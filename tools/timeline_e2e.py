"""Per-launch timeline of one END-TO-END proof (ProverState.prove: host synthesis, H2D, proof; k = 19 by default), same CSV as
tools/timeline.py (name,stream,start_ms,duration_ms).  Development aid: shows where the host-side phases leave the device idle."""
import importlib, os, sys, time
sys.path.insert(0, os.getcwd())
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline_e2e.csv"
degree = int(sys.argv[2]) if len(sys.argv) > 2 else 19
if os.path.exists(out):
    os.remove(out)
zkw = importlib.import_module("webauthn-halo2_b200")
st = zkw.ProverState(zkw.CircuitParams.for_degree(degree), 0)
a = [zkw.synthetic_assertion(i) for i in range(4)]
for i in range(3): st.prove(a[i], zkw.TRANSCRIPT_EVM)
st.ctx.sync()
os.environ["ZKW_TIMELINE"] = out
st.ctx.profile_enable(True)
st.ctx.profile_reset()
t0 = time.perf_counter()
st.prove(a[3], zkw.TRANSCRIPT_EVM)
wall = (time.perf_counter() - t0) * 1e3
st.ctx.profile_all()
st.ctx.profile_enable(False)
rows = [l.strip().split(",") for l in open(out)]
t_end = max(float(r[2]) + float(r[3]) for r in rows)
print(f"{len(rows)} launches, device span {t_end:.2f} ms, wall {wall:.2f} ms, host synthesis {st.last_synthesis_ms if hasattr(st, 'last_synthesis_ms') else -1}")
for r in rows[:40]:
    print(f"{r[0][:32]:32s} {r[1][-5:]} {float(r[2]):7.3f} {float(r[3]):7.3f}")
