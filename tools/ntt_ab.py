import importlib, sys, os
import numpy as np, torch
sys.path.insert(0, os.getcwd())
zkw = importlib.import_module("webauthn-halo2_b200")
ctx = zkw.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=0); torch.cuda.set_stream(stream)
def timed(fn, reps=20):
    fn(); stream.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps): fn()
    b.record(stream); b.synchronize()
    return a.elapsed_time(b)/reps
out=[]
for k in (19,21):
    n=1<<k
    a=torch.randint(0,1<<62,(n,4),dtype=torch.int64,device="cuda"); a[:,3]&=(1<<60)-1
    out.append("k=%d intt %.4f ms"%(k,timed(lambda: ctx.lagrange_to_coeff_dev(a,k))))
k=19; n=1<<k
a=torch.randint(0,1<<62,(n,4),dtype=torch.int64,device="cuda"); a[:,3]&=(1<<60)-1
e=torch.empty((4*n,4),dtype=torch.int64,device="cuda")
out.append("coset 19->21 %.4f ms"%timed(lambda: ctx.coeff_to_extended_dev(a,k,k+2,e)))
print(os.environ.get("ZKW_B200_LIB","default")[-12:], " | ".join(out))
