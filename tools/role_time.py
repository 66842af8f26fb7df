import os, sys, json
ROOT="/root/repo"; sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "steroscopic-hardware_b200"))
import numpy as np, torch, despair, ctypes
from despair import _native as N
ctx = despair.Context([0], 1920, 1080, 1)
Hh, Ww, B, D, F = 1080, 1920, 9, 128, 16
rng = np.random.default_rng(1)
L = torch.from_numpy(rng.integers(0,256,(F,Hh,Ww),dtype=np.uint8)).cuda(); R = torch.roll(L, -20, 2).contiguous(); O = torch.zeros_like(L)
st = torch.cuda.current_stream().cuda_stream
for skip in (0, 1, 2, 6, 5, 4):
    t = N.Tuning(); t.kernel_variant = 3; t.reserved[1] = skip
    run = lambda: N.check(N.lib().sadgpu_compute_device_batch(ctx._h, 0, F, L.data_ptr(), Ww, Ww*Hh, R.data_ptr(), Ww, Ww*Hh, Ww, Hh, B, D, 0, Hh, O.data_ptr(), Ww, Ww*Hh, st, ctypes.byref(t)))
    for _ in range(2): run()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): run()
    e1.record(); torch.cuda.synchronize()
    print("debug_skip", skip, f"{e0.elapsed_time(e1)/5/F*1e3:.1f} us/frame")
    if skip & 4:
        buf = (ctypes.c_uint32 * 96)()
        N.check(N.lib().sadgpu_debug_read(ctx._h, 0, buf, 96))
        v = np.array(buf[:]).reshape(24, 4); nb = v[0, 3]
        print("  cycles/batch [work, commit, barrier-wait] by warp:", {w: (v[w, :3] / max(nb + 2, 1)).round(0).astype(int).tolist() for w in range(24) if v[w, 3]})

