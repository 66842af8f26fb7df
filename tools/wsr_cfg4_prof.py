"""One launch of the large-window kernel at BASELINE configs[3] (3840x2160, block 31, max disparity 256, 2 frames) for an ncu capture;
prints the CUDA-event time when run plain."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "steroscopic-hardware_b200"))
import numpy as np, torch, despair
W, H, F, B, D = 3840, 2160, 2, 31, 256
ctx = despair.Context([0], W, H, 1)
rng = np.random.default_rng(1)
L = torch.from_numpy(rng.integers(0, 256, (F, H, W), dtype=np.uint8)).cuda(); R = torch.roll(L, -20, 2).contiguous(); O = torch.zeros_like(L)
st = torch.cuda.current_stream().cuda_stream
run = lambda: ctx.compute_device_batch(F, L.data_ptr(), W, W * H, R.data_ptr(), W, W * H, W, H, B, D, O.data_ptr(), W, W * H, cuda_stream=st)
run(); torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); run(); run(); e1.record(); torch.cuda.synchronize()
print(f"cfg4: {e0.elapsed_time(e1) / F / 2 * 1e3:.1f} us/frame", despair.plan_describe(W, H, B, D, frames=F))
