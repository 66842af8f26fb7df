"""Per-role wait accounting of the vertical-first kernel (library built with -DVH_PROFILE).  env: SADGPU_LIB, B, D, F"""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "steroscopic-hardware_b200"))
import numpy as np, torch
from despair import _native as N
if os.environ.get("SADGPU_LIB"): N.LIB_PATH = os.environ["SADGPU_LIB"]
import despair
B, D, F = (int(os.environ.get(k, d)) for k, d in (("B", 15), ("D", 128), ("F", 8)))
W, H = 1920, 1080
ctx = despair.Context([0], W, H, 1)
rng = np.random.default_rng(1)
L = torch.from_numpy(rng.integers(0, 256, (F, H, W), dtype=np.uint8)).cuda(); R = torch.roll(L, -20, 2).contiguous(); O = torch.zeros_like(L)
st = torch.cuda.current_stream().cuda_stream
for flags in (0, 4):
    t = N.Tuning(); t.kernel_variant = int(os.environ.get("VARIANT", 6)); t.reserved[1] = flags
    run = lambda: N.check(N.lib().sadgpu_compute_device_batch(ctx._h, 0, F, L.data_ptr(), W, W * H, R.data_ptr(), W, W * H, W, H, B, D, 0, H, O.data_ptr(), W, W * H, st, ctypes.byref(t)))
    for _ in range(2): run()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): run()
    e1.record(); torch.cuda.synchronize()
    print(f"flags {flags}: {e0.elapsed_time(e1) / 5 / F * 1e3:.1f} us/frame", despair.plan_describe(W, H, B, D, tuning=dict(kernel_variant=int(os.environ.get("VARIANT", 6))), frames=F)["grid"])
buf = (ctypes.c_uint32 * 192)()
N.check(N.lib().sadgpu_debug_read(ctx._h, 0, buf, 192))
v = np.array(buf[:], dtype=np.int64).reshape(24, 8) * 16
names = ["tile_full", "c_empty|h_empty", "tile_empty", "c_full|h_full", "tail_full|pk_full", "tail_empty|pk_empty", "-", "total"]
for w in range(24):
    if v[w, 7]:
        print(f"warp {w:2d} total {v[w,7]:9d} clk  " + "  ".join(f"{names[i]} {100 * v[w, i] / v[w, 7]:5.1f}%" for i in range(6) if v[w, i]))
