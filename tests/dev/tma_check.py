import os, sys, ctypes
ROOT="/root/repo"; sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "steroscopic-hardware_b200"))
import numpy as np, torch, despair
from despair import _native as N
from oracle import oracle as O
ctx = despair.Context([0], 2048, 1200, 1)
rng = np.random.default_rng(3)
st = torch.cuda.current_stream().cuda_stream
bad = 0
for (W, H, B, D, F) in [(64, 20, 9, 128, 1), (128, 37, 9, 128, 2), (320, 50, 5, 200, 3), (1920, 64, 9, 128, 2), (96, 9, 1, 68, 1), (160, 33, 8, 256, 2), (48, 70, 9, 100, 1)]:
    Ls = rng.integers(0,256,(F,H,W),dtype=np.uint8); Rs = rng.integers(0,256,(F,H,W),dtype=np.uint8)
    dL = torch.from_numpy(Ls).cuda(); dR = torch.from_numpy(Rs).cuda()
    for notma in (0, 1):
        dO = torch.zeros_like(dL)
        t = N.Tuning(); t.kernel_variant = 3; t.reserved[2] = notma
        N.check(N.lib().sadgpu_compute_device_batch(ctx._h, 0, F, dL.data_ptr(), W, W*H, dR.data_ptr(), W, W*H, W, H, B, D, 0, H, dO.data_ptr(), W, W*H, st, ctypes.byref(t)))
        torch.cuda.synchronize()
        got = dO.cpu().numpy()
        ok = all(np.array_equal(got[f], O.frame_box(Ls[f], Rs[f], B, D)) for f in range(F))
        bad += (not ok)
        print(f"W={W} H={H} B={B} D={D} F={F} notma={notma}: {'OK' if ok else 'MISMATCH'}", flush=True)
print("bad", bad)
