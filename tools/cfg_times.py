"""Device-resident kernel times of the BASELINE.json configs (CUDA events, >=3 warm-up, batch launches where
the kernel supports them) -> gpurun_out/cfg_times.json.   python tools/cfg_times.py [--sweep]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "steroscopic-hardware_b200"))
import numpy as np, torch, despair

P_INT = 63.9 * 148 * 1.965e9 / 1e12

L2_BYTES = 126 << 20

def time_cfg(ctx, W, H, B, D, F, reps=6):
    """Launches rotate over enough distinct input / output sets that every launch streams from HBM (the sets together exceed
    twice the 126 MB L2), as bench.py does with its 256-frame steps."""
    rng = np.random.default_rng(1)
    nset = max(2, -(-2 * L2_BYTES // (3 * F * W * H)))
    base = torch.from_numpy(rng.integers(0, 256, (H, W), dtype=np.uint8)).cuda()
    L = base.unsqueeze(0).repeat(nset * F, 1, 1).contiguous(); R = torch.roll(L, -20, 2).contiguous(); O = torch.zeros_like(L)
    st = torch.cuda.current_stream().cuda_stream
    def run(k):
        o = (k % nset) * F
        ctx.compute_device_batch(F, L[o].data_ptr(), W, W * H, R[o].data_ptr(), W, W * H, W, H, B, D, O[o].data_ptr(), W, W * H, cuda_stream=st)
    for k in range(3): run(k)
    reps = max(reps, nset)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for k in range(reps): run(k)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * F)
    ev = W * H * (D + 1)
    return {"W": W, "H": H, "B": B, "D": D, "frames_per_launch": F, "us_per_frame": round(us, 1), "frames_per_s": round(1e6 / us, 1),
            "Mpix_D_per_s": round(W * H * D / us, 1), "int_alu_roofline_frac": round(6 * ev / (us * 1e-6) / 1e12 / P_INT, 3),
            "variant": despair.plan_describe(W, H, B, D, frames=F)["variant"]}

def main():
    ctx = despair.Context([0], 3840, 2160, 1)
    out = {"configs": [], "sweep": []}
    for name, (W, H, B, D, F) in {"cfg1 640x480 B9 D64": (640, 480, 9, 64, 64), "cfg2 1080p B15 D256": (1920, 1080, 15, 256, 16),
                                  "cfg3 1080p B9 D128": (1920, 1080, 9, 128, 16), "cfg4 4K B31 D256 (one GPU)": (3840, 2160, 31, 256, 8)}.items():
        r = time_cfg(ctx, W, H, B, D, F); r["name"] = name; out["configs"].append(r); print(r, flush=True)
    if "--sweep" in sys.argv:
        for B in (3, 5, 7, 9, 11, 13, 15, 16, 17, 21, 25, 31):
            for D in (16, 32, 64, 128, 256):
                r = time_cfg(ctx, 1920, 1080, B, D, 8, reps=3); out["sweep"].append(r); print(r, flush=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "cfg_times.json"), "w"), indent=1)

main()
