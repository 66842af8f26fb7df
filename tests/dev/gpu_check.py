"""Developer battery: GPU path vs oracle on many small cases + the fixtures.  Run on the GPU box:
   python tools/gpu_check.py [--quick]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "steroscopic-hardware_b200"))
import numpy as np
import torch
from oracle import oracle as O
import despair

def dev_run(ctx, L, R, B, D, tuning=None, y0=0, y1=None):
    h, w = L.shape
    dL = torch.from_numpy(L).cuda(); dR = torch.from_numpy(R).cuda()
    dO = torch.full((h, w), 77, dtype=torch.uint8, device="cuda")
    ctx.compute_device(dL.data_ptr(), w, dR.data_ptr(), w, w, h, B, D, dO.data_ptr(), w, y0=y0, y1=y1,
                       cuda_stream=torch.cuda.current_stream().cuda_stream, tuning=tuning)
    torch.cuda.synchronize()
    return dO.cpu().numpy()

def main():
    O.build()
    ctx = despair.Context([0], 4096, 2304, 2)
    rng = np.random.default_rng(11)
    nbad = 0; ncase = 0
    t0 = time.time()
    for i in range(120 if "--quick" not in sys.argv else 30):
        W = int(rng.integers(1, 200)); H = int(rng.integers(1, 90))
        B = int(rng.integers(1, 32)); D = int(rng.choice([1, 3, 5, 16, 17, 31, 64, 100, 128, 200, 255, 256]))
        kind = i % 4
        if kind == 0: L = rng.integers(0, 256, (H, W), dtype=np.uint8); R = rng.integers(0, 256, (H, W), dtype=np.uint8)
        elif kind == 1:
            base = rng.integers(0, 256, (H, W + 40), dtype=np.uint8); s = int(rng.integers(0, 30))
            L = base[:, 40:40 + W].copy(); R = base[:, 40 - s:40 - s + W].copy() if False else np.roll(base, -s, 1)[:, 40:40 + W].copy()
        elif kind == 2: L = rng.integers(0, 3, (H, W), dtype=np.uint8); R = rng.integers(0, 3, (H, W), dtype=np.uint8)
        else: L = np.full((H, W), 255, np.uint8); R = np.zeros((H, W), np.uint8)
        exp = O.frame_box(L, R, B, D)
        tun = None
        if i % 3 == 1: tun = dict(rows_per_batch=int(rng.integers(1, 9)), band_rows=int(rng.integers(1, 40)), groups_per_chunk=int(rng.integers(1, 21)))
        if B <= 15 and i % 2 == 0:
            tun = dict(tun or {}); tun['kernel_variant'] = 2
        if B <= 9 and D >= 68 and i % 4 == 1:
            tun = dict(tun or {}); tun['kernel_variant'] = 3
        if B >= 16 and i % 2 == 0:
            tun = dict(band_rows=(tun or {}).get('band_rows', 0), kernel_variant=4)
        if B >= 16 and i % 2 == 1:
            tun = dict(tun or {}); tun['kernel_variant'] = 1
        got = dev_run(ctx, L, R, B, D, tun)
        ncase += 1
        if not np.array_equal(got, exp):
            nbad += 1
            ys, xs = np.nonzero(got != exp)
            print(f"MISMATCH W={W} H={H} B={B} D={D} kind={kind} tun={tun} n={len(ys)} first=({xs[0]},{ys[0]}) got={got[ys[0], xs[0]]} exp={exp[ys[0], xs[0]]} xr=({xs.min()},{xs.max()}) yr=({ys.min()},{ys.max()})")
    print(f"random: {ncase} cases, {nbad} bad, {time.time() - t0:.1f}s")
    # host API + fixtures
    from PIL import Image
    G = os.path.join(ROOT, "tests", "golden")
    ld = lambda n: np.array(Image.open(os.path.join(G, n)), np.uint8)
    for tag in ("00001", "00002", "00335", "01000"):
        L = ld(f"L_{tag}_gray.png"); R = ld(f"R_{tag}_gray.png"); exp = ld(f"disp_{tag}_b9_d64.png")
        got = ctx.compute(L, R, 9, 64)
        print("fixture", tag, "OK" if np.array_equal(got, exp) else f"BAD {(got != exp).sum()}")
    L = ld("im0_intended_gray.png"); R = ld("im1_intended_gray.png"); exp = ld("disp_im0_im1_intended_b15_d256.png")
    t = time.time(); got = ctx.compute(L, R, 15, 256); dt = time.time() - t
    print("cfg2", "OK" if np.array_equal(got, exp) else f"BAD {(got != exp).sum()}", f"{dt*1e3:.2f} ms host call")
    # timing cfg3
    Hh, Ww = 1080, 1920
    Ls = rng.integers(0, 256, (Hh, Ww), dtype=np.uint8); Rs = np.roll(Ls, -20, 1)
    dL = torch.from_numpy(Ls).cuda(); dR = torch.from_numpy(Rs).cuda(); dO = torch.zeros((Hh, Ww), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for (B, D) in [(9, 128), (15, 256), (31, 256), (9, 64), (3, 16)]:
        for _ in range(3): ctx.compute_device(dL.data_ptr(), Ww, dR.data_ptr(), Ww, Ww, Hh, B, D, dO.data_ptr(), Ww, cuda_stream=st)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 20
        for _ in range(n): ctx.compute_device(dL.data_ptr(), Ww, dR.data_ptr(), Ww, Ww, Hh, B, D, dO.data_ptr(), Ww, cuda_stream=st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        ev = Ww * Hh * (D + 1)
        print(f"1080p B={B} D={D}: {ms*1e3:.1f} us/frame  {ev/ms/1e9:.3f} Tevals/s  plan={json.dumps(despair.plan_describe(Ww, Hh, B, D))}")
        exp = O.frame_box(Ls, Rs, B, D, 500, 516)
        print("   parity rows 500..516:", np.array_equal(dO.cpu().numpy()[500:516], exp))
        if B <= 15:
            F = 16
            bL = dL.unsqueeze(0).repeat(F, 1, 1).contiguous(); bR = dR.unsqueeze(0).repeat(F, 1, 1).contiguous(); bO = torch.zeros_like(bL)
            fs = Hh * Ww
            run = lambda: ctx.compute_device_batch(F, bL.data_ptr(), Ww, fs, bR.data_ptr(), Ww, fs, Ww, Hh, B, D, bO.data_ptr(), Ww, fs, cuda_stream=st)
            for _ in range(2): run()
            e0.record()
            for _ in range(5): run()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / (5 * F)
            print(f"   batch16: {ms*1e3:.1f} us/frame  {ev/ms/1e9:.3f} Tevals/s  plan={json.dumps(despair.plan_describe(Ww, Hh, B, D, frames=F))}")
            ok = all(np.array_equal(bO[f].cpu().numpy()[500:516], exp) for f in (0, F - 1))
            print("   batch parity:", ok, " full-frame equal to single:", bool(torch.equal(bO[3], dO)))

if __name__ == "__main__":
    main()
