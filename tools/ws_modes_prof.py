"""Launches the warp-specialised kernel once per lane layout (max disparity 128 / 64 / 32 / 16, block 9, 8 frames of 1080p) for an ncu
capture; prints CUDA-event times when run plain.  env: B (default 9), DS (comma list)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "steroscopic-hardware_b200"))
import numpy as np, torch
from despair import _native as N
if os.environ.get("SADGPU_LIB"): N.LIB_PATH = os.environ["SADGPU_LIB"]
import despair
BS = [int(x) for x in os.environ.get("BS", os.environ.get("B", "9")).split(",")]; DS = [int(x) for x in os.environ.get("DS", "128,64,32,16").split(",")]
VAR = int(os.environ.get("VARIANT", 0)); TUN = dict(kernel_variant=VAR) if VAR else None
if os.environ.get("GPC"): TUN = dict(TUN or {}, groups_per_chunk=int(os.environ["GPC"]))
W, H, F = 1920, 1080, 8
ctx = despair.Context([0], W, H, 1)
rng = np.random.default_rng(1)
for B in BS:
    if os.environ.get("CHECK"):            # developer A/B runs: a small parity check against the oracle first
        from oracle import oracle as O
        bad = 0
        for D in DS:
            for (hh, ww) in ((37, 208), (64, 336), (37, 200), (45, 32)):
                l = rng.integers(0, 256, (hh, ww), dtype=np.uint8); r = np.roll(l, -7, 1)
                dl = torch.from_numpy(l).cuda(); dr = torch.from_numpy(r).cuda(); do = torch.zeros_like(dl)
                ctx.compute_device(dl.data_ptr(), ww, dr.data_ptr(), ww, ww, hh, B, D, do.data_ptr(), ww, cuda_stream=torch.cuda.current_stream().cuda_stream, tuning=TUN)
                torch.cuda.synchronize()
                bad += int((do.cpu().numpy() != O.frame_box(l, r, B, D)).sum())
        print("parity check: wrong pixels =", bad, flush=True)
    L = torch.from_numpy(rng.integers(0, 256, (F, H, W), dtype=np.uint8)).cuda(); R = torch.roll(L, -20, 2).contiguous(); O = torch.zeros_like(L)
    st = torch.cuda.current_stream().cuda_stream
    for D in DS:
        run = lambda: ctx.compute_device_batch(F, L.data_ptr(), W, W * H, R.data_ptr(), W, W * H, W, H, B, D, O.data_ptr(), W, W * H, cuda_stream=st, tuning=TUN)
        run(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); run(); run(); run(); e1.record(); torch.cuda.synchronize()
        print(f"B={B} D={D}: {e0.elapsed_time(e1) / F / 3 * 1e3:.1f} us/frame", despair.plan_describe(W, H, B, D, frames=F, tuning=TUN), flush=True)
