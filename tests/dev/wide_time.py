import os, sys
ROOT="/root/repo"; sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "steroscopic-hardware_b200"))
import numpy as np, torch, despair
from oracle import oracle as O
ctx = despair.Context([0], 1920, 1080, 1)
Hh, Ww, F = 1080, 1920, 4
rng = np.random.default_rng(1)
Ln = rng.integers(0,256,(Hh,Ww),dtype=np.uint8); Rn = np.roll(Ln,-20,1)
L = torch.from_numpy(Ln).cuda().unsqueeze(0).repeat(F,1,1).contiguous(); R = torch.from_numpy(Rn).cuda().unsqueeze(0).repeat(F,1,1).contiguous(); Oo = torch.zeros_like(L)
st = torch.cuda.current_stream().cuda_stream
for (B, D) in [(15,256),(15,128),(15,64),(13,256),(11,256),(11,128),(11,64)]:
    run = lambda: ctx.compute_device_batch(F, L.data_ptr(), Ww, Ww*Hh, R.data_ptr(), Ww, Ww*Hh, Ww, Hh, B, D, Oo.data_ptr(), Ww, Ww*Hh, cuda_stream=st)
    for _ in range(2): run()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1)/3/F*1e3
    ok = np.array_equal(Oo[F-1].cpu().numpy()[300:308], O.frame_box(Ln, Rn, B, D, 300, 308))
    print(f"B={B} D={D}: {us:.1f} us/frame  frac={6*Ww*Hh*(D+1)/(us*1e-6)/1e12/18.586:.3f} parity={ok}", flush=True)
