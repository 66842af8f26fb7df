"""Developer soak: random (size, block, max disparity, tiling, chunk size, row range, frames) through the planner's kernels against the
oracle for a fixed wall time (block >= BMIN, default 16).  BMIN=1 python tests/dev/soak_planner_kernels.py [seconds] [seed]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "steroscopic-hardware_b200"))
import numpy as np, torch, despair
from oracle import oracle as O

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
ctx = despair.Context([0], 1024, 512, 2)
st = torch.cuda.current_stream().cuda_stream
t0 = time.time(); n = 0; bad = 0
while time.time() - t0 < budget:
    B = int(rng.integers(int(os.environ.get("BMIN", 16)), 32)); D = int(rng.choice([int(rng.integers(1, 257)), 16, 32, 48, 64, 128, 255, 256]))
    W = 16 * int(rng.integers(1, 40)) if rng.random() < 0.8 else int(rng.integers(1, 600)); H = int(rng.integers(1, 200)); F = int(rng.integers(1, 4))
    kind = int(rng.integers(0, 4))
    if kind == 0: L = rng.integers(0, 256, (F, H, W), dtype=np.uint8); R = rng.integers(0, 256, (F, H, W), dtype=np.uint8)
    elif kind == 1: L = rng.integers(0, 256, (F, H, W), dtype=np.uint8); R = np.roll(L, -int(rng.integers(0, 40)), 2).copy()
    elif kind == 2: L = rng.integers(0, 3, (F, H, W), dtype=np.uint8); R = rng.integers(0, 3, (F, H, W), dtype=np.uint8)
    else: L = np.full((F, H, W), 255, np.uint8); R = np.zeros((F, H, W), np.uint8)
    tun = {}
    if rng.random() < 0.5: tun["band_rows"] = int(rng.integers(1, 60))
    if rng.random() < 0.5: tun["groups_per_chunk"] = int(rng.choice([5, 9, 13, 17, 33]))
    y0 = int(rng.integers(0, H)) if rng.random() < 0.3 else 0; y1 = int(rng.integers(y0, H + 1)) if rng.random() < 0.3 else H
    dL = torch.from_numpy(L).cuda(); dR = torch.from_numpy(R).cuda(); dO = torch.full_like(dL, 77)
    ctx.compute_device_batch(F, dL.data_ptr(), W, W * H, dR.data_ptr(), W, W * H, W, H, B, D, dO.data_ptr(), W, W * H, y0=y0, y1=y1,
                             cuda_stream=st, tuning=tun or None)
    torch.cuda.synchronize()
    got = dO.cpu().numpy()
    for f in range(F):
        exp = O.frame_box(L[f], R[f], B, D, y0, y1)
        if not (np.array_equal(got[f, y0:y1], exp) and (got[f, :y0] == 77).all() and (got[f, y1:] == 77).all()):
            bad += 1; print("MISMATCH", dict(W=W, H=H, F=F, B=B, D=D, y0=y0, y1=y1, tun=tun, kind=kind, f=f), flush=True)
    n += 1
print(f"soak: {n} random configurations, {bad} mismatches, {time.time() - t0:.0f} s")
