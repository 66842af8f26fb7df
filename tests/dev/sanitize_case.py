"""Small invocation of every kernel variant for compute-sanitizer (memcheck / racecheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "steroscopic-hardware_b200"))
import numpy as np, torch, despair
from oracle import oracle as O
rng = np.random.default_rng(2)
ctx = despair.Context([0], 512, 256, 1)
ok = True
for (W, H, B, D, var) in [(97, 41, 9, 128, 3), (130, 37, 15, 256, 2), (75, 29, 31, 64, 1), (64, 20, 3, 16, 2), (200, 30, 8, 200, 3)]:
    L = rng.integers(0, 256, (H, W), dtype=np.uint8); R = rng.integers(0, 256, (H, W), dtype=np.uint8)
    dL = torch.from_numpy(L).cuda(); dR = torch.from_numpy(R).cuda(); dO = torch.zeros_like(dL)
    ctx.compute_device(dL.data_ptr(), W, dR.data_ptr(), W, W, H, B, D, dO.data_ptr(), W, cuda_stream=torch.cuda.current_stream().cuda_stream, tuning=dict(kernel_variant=var))
    torch.cuda.synchronize()
    ok &= bool(np.array_equal(dO.cpu().numpy(), O.frame_box(L, R, B, D)))
print("sanitize_case parity:", ok)
