"""Times the cfg3 batch launch (16 frames, 1080p, B=9, D=128) of the library named by SADGPU_LIB (developer A/B tool)."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "steroscopic-hardware_b200"))
import numpy as np, torch
from despair import _native as N
if os.environ.get("SADGPU_LIB"):
    N.LIB_PATH = os.environ["SADGPU_LIB"]
import despair
B, D = int(os.environ.get("B", 9)), int(os.environ.get("D", 128))
VAR = int(os.environ.get("VARIANT", 0))
ctx = despair.Context([0], 1920, 1080, 1)
Hh, Ww, F = 1080, 1920, 16
rng = np.random.default_rng(1)
L = torch.from_numpy(rng.integers(0, 256, (F, Hh, Ww), dtype=np.uint8)).cuda(); R = torch.roll(L, -20, 2).contiguous(); O = torch.zeros_like(L)
st = torch.cuda.current_stream().cuda_stream
run = lambda: ctx.compute_device_batch(F, L.data_ptr(), Ww, Ww * Hh, R.data_ptr(), Ww, Ww * Hh, Ww, Hh, B, D, O.data_ptr(), Ww, Ww * Hh, cuda_stream=st, tuning=dict(kernel_variant=VAR) if VAR else None)
for _ in range(3): run()
best = 1e9
for rep in range(3):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 10 / F * 1e3)
print(os.environ.get("SADGPU_LIB", "default"), f"B={B} D={D}: {best:.2f} us/frame", despair.plan_describe(Ww, Hh, B, D, frames=F)["variant"], "sum", int(O.sum().item()))
