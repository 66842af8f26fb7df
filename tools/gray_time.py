"""N1: HBM roofline of the Go-exact luma kernel sadgpu_gray_device.  Frames are cycled through a set larger
than L2 (126 MB) so that every launch streams from HBM.  -> gpurun_out/gray_times.json"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "steroscopic-hardware_b200"))
import numpy as np, torch, despair
peak = 6549.8
try: peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception: pass
ctx = despair.Context([0], 3840, 2160, 1)
st = torch.cuda.current_stream().cuda_stream
out = []
for (W, H, ch, mode, name) in ((1920, 1080, 4, 0, "NRGBA8 1080p (testdata-style RGBA PNG)"), (3840, 2160, 4, 0, "NRGBA8 4K"),
                               (1920, 1080, 3, 1, "RGB8 intended luma 1080p (im0/im1-style RGB PNG)")):
    # one launch over a stack of frames (a tall image of nset*H rows, < 65 536 grid rows): kernel time, not launch overhead
    nset = max(2, min(65535 // H, int(400e6 // (W * H * ch))))
    src = torch.randint(0, 256, (2, nset * H, W * ch), dtype=torch.uint8, device="cuda")
    if ch == 4:                                          # like the testdata PNGs: opaque except for a sprinkle of alpha 251..254
        a = src.view(2, nset * H, W, 4)[..., 3]
        a.copy_(torch.where(torch.rand(a.shape, device="cuda") < 0.01, torch.randint(251, 255, a.shape, dtype=torch.uint8, device="cuda"), torch.full_like(a, 255)))
    dst = torch.empty((2, nset * H, W), dtype=torch.uint8, device="cuda")
    run = lambda k: ctx.gray_device(src[k].data_ptr(), W * ch, ch, mode, W, nset * H, dst[k].data_ptr(), W, cuda_stream=st)
    for k in range(2): run(k)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    reps = 5
    for _ in range(reps):
        for k in range(2): run(k)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * 2 * nset)
    gbs = W * H * (ch + 1) / us / 1e3
    r = {"case": name, "us_per_frame": round(us, 2), "algorithmic_bytes": W * H * (ch + 1), "GB_per_s": round(gbs, 1), "hbm_peak_GB_per_s": peak, "frac": round(gbs / peak, 3)}
    out.append(r); print(r, flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "gray_times.json"), "w"), indent=1)
