import os, sys, time
ROOT="/root/repo"; sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "steroscopic-hardware_b200"))
import numpy as np, torch, despair
W,H,B,D=1920,1080,9,128
rng=np.random.default_rng(0)
for ns in (2,4,8):
    ctx=despair.Context([0],W,H,ns)
    pin=[ctx.host_pair(H,W) for _ in range(4)]
    for a,b in pin: a[:]=rng.integers(0,256,(H,W),dtype=np.uint8); b[:]=np.roll(a,-20,1)
    outs=[ctx.host_array((H,W)) for _ in range(ns)]
    def run(F, into):
        t=[None]*ns
        for k in range(F):
            s=k%ns
            if t[s] is not None: ctx.wait(t[s],None if into else outs[s])
            t[s]=ctx.submit(pin[k%4][0],pin[k%4][1],B,D,stream=s,out=outs[s] if into else None)
        for s in range(ns):
            if t[s] is not None: ctx.wait(t[s],None if into else outs[s])
    for into in (False, True):
        run(16, into); torch.cuda.synchronize()
        t0=time.perf_counter(); run(512, into); dt=time.perf_counter()-t0
        print(f"streams={ns} submit_into={into}: {512/dt:.0f} fps  {dt/512*1e6:.1f} us/frame", flush=True)
    ctx.close()
