"""Developer check of the vertical-first kernel (kernel_variant=5) against the oracle + timing vs the default plan."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "steroscopic-hardware_b200"))
import numpy as np, torch
from oracle import oracle as O
from despair import _native as N
if os.environ.get("SADGPU_LIB"): N.LIB_PATH = os.environ["SADGPU_LIB"]
import despair
VAR = int(os.environ.get("VARIANT", 5))
BS = [int(b) for b in os.environ.get("BS", "11,13,15,12,17,21,31").split(",")]

def dev_run(ctx, L, R, B, D, tuning=None):
    h, w = L.shape
    dL = torch.from_numpy(L).cuda(); dR = torch.from_numpy(R).cuda()
    dO = torch.full((h, w), 77, dtype=torch.uint8, device="cuda")
    ctx.compute_device(dL.data_ptr(), w, dR.data_ptr(), w, w, h, B, D, dO.data_ptr(), w,
                       cuda_stream=torch.cuda.current_stream().cuda_stream, tuning=tuning)
    torch.cuda.synchronize()
    return dO.cpu().numpy()

def timed(ctx, W, H, B, D, F, tuning):
    rng = np.random.default_rng(1)
    base = torch.from_numpy(rng.integers(0, 256, (H, W), dtype=np.uint8)).cuda()
    L = base.unsqueeze(0).repeat(F, 1, 1).contiguous(); R = torch.roll(L, -20, 2).contiguous(); Oo = torch.zeros_like(L)
    st = torch.cuda.current_stream().cuda_stream
    run = lambda: ctx.compute_device_batch(F, L.data_ptr(), W, W * H, R.data_ptr(), W, W * H, W, H, B, D, Oo.data_ptr(), W, W * H, cuda_stream=st, tuning=tuning)
    for _ in range(2): run()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(3): run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (3 * F), int(Oo.sum().item())

def main():
    O.build()
    ctx = despair.Context([0], 3840, 2160, 1)
    rng = np.random.default_rng(5)
    nbad = 0
    cases = []
    for B in BS:
        for (W, H, D) in ((150, 40, 128), (97, 23, 64), (333, 70, 256), (64, 9, 16), (200, 33, 200)):
            cases.append((W, H, B, D))
    quick = "--quick" in sys.argv
    for i, (W, H, B, D) in enumerate(cases[:6] if quick else cases):
        kind = i % 3
        if kind == 0: L = rng.integers(0, 256, (H, W), dtype=np.uint8); R = rng.integers(0, 256, (H, W), dtype=np.uint8)
        elif kind == 1:
            base = rng.integers(0, 256, (H, W + 40), dtype=np.uint8); s = int(rng.integers(0, 30))
            L = base[:, 40:40 + W].copy(); R = np.roll(base, -s, 1)[:, 40:40 + W].copy()
        else: L = np.full((H, W), 255, np.uint8); R = np.zeros((H, W), np.uint8)
        exp = O.frame_box(L, R, B, D)
        for tun in (dict(kernel_variant=VAR), dict(kernel_variant=VAR, band_rows=int(rng.integers(3, 30)))):
            got = dev_run(ctx, L, R, B, D, tun)
            if not np.array_equal(got, exp):
                nbad += 1
                ys, xs = np.nonzero(got != exp)
                print(f"MISMATCH W={W} H={H} B={B} D={D} kind={kind} tun={tun} n={len(ys)} first=({xs[0]},{ys[0]}) got={got[ys[0], xs[0]]} exp={exp[ys[0], xs[0]]} xr=({xs.min()},{xs.max()}) yr=({ys.min()},{ys.max()})", flush=True)
    print(f"vh cases: {len(cases) * 2}, bad {nbad}", flush=True)
    if "--notime" in sys.argv: return
    for (W, H, B, D, F) in ((1920, 1080, 15, 256, 8), (1920, 1080, 11, 128, 8), (1920, 1080, 13, 128, 8), (1920, 1080, 15, 128, 8),
                            (1920, 1080, 17, 128, 8), (1920, 1080, 21, 256, 4), (3840, 2160, 31, 256, 2)):
        if B not in BS: continue
        t0, s0 = timed(ctx, W, H, B, D, F, None)
        t5, s5 = timed(ctx, W, H, B, D, F, dict(kernel_variant=VAR))
        roof = 6 * W * H * (D + 1) / 18.5863e12 * 1e6
        print(f"{W}x{H} B={B} D={D}: default {t0:.1f} us/frame ({roof / t0:.3f})   variant {t5:.1f} us/frame ({roof / t5:.3f})   same={s0 == s5}", flush=True)

main()
