"""One batch launch of a forced kernel variant (for ncu captures).  env: B, D, F, VARIANT"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "steroscopic-hardware_b200"))
import numpy as np, torch
from despair import _native as N
if os.environ.get("SADGPU_LIB"): N.LIB_PATH = os.environ["SADGPU_LIB"]
import despair
B, D, F, V = (int(os.environ.get(k, d)) for k, d in (("B", 15), ("D", 128), ("F", 8), ("VARIANT", 5)))
W, H = 1920, 1080
ctx = despair.Context([0], W, H, 1)
rng = np.random.default_rng(1)
L = torch.from_numpy(rng.integers(0, 256, (F, H, W), dtype=np.uint8)).cuda(); R = torch.roll(L, -20, 2).contiguous(); O = torch.zeros_like(L)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    ctx.compute_device_batch(F, L.data_ptr(), W, W * H, R.data_ptr(), W, W * H, W, H, B, D, O.data_ptr(), W, W * H, cuda_stream=st,
                             tuning=dict(kernel_variant=V) if V else None)
torch.cuda.synchronize()
print("ok", int(O.sum().item()))
