#!/usr/bin/env python
"""bench.py — headline benchmark of the SAD block-matching disparity path on B200.

Workload (BASELINE.json metric, configs[2] / SURVEY.md §8(d) cfg3): synthetic 1920x1080 grayscale
stereo stream, block 9, max disparity 128.  One *step* = one pass of the hot path over a batch of
FRAMES frame pairs per GPU (inputs 2*FRAMES*W*H bytes > the 126 MB L2, so every step streams from
HBM).  Frame-sharded over ranks, no collective on the data path: weak scaling.

  value        whole-job Mpix*D/s, inputs resident in HBM, CUDA events on the launch stream
  e2e          same metric through the C ABI (sadgpu_submit / sadgpu_wait) with host buffers:
               pinned H2D of both images and D2H of the map inside the timed region
  roofline     integer-ALU roofline (the binding one): 6 int-ops per evaluation
               against the measured 64 lane-ops/clk/SM; HBM figures beside it
  cpu_baseline the oracle's literal restatement of pkg/despair on the host cores (rank 0, N=1)

  extra        every other BASELINE.json config, measured and parity-checked in the same run: cfg1, cfg2 (intended luma and the
               all-zero images LoadPNG really produces), cfg4 on one GPU and row-band sharded over all N GPUs
               (sadgpu_compute_sharded), the cfg5 sweep (1024 pairs frame-sharded over the ranks x the full WebUI grid
               B {3..31 odd, 16} x D {16..256 step 16}), and the UNCHANGED reference call pattern of OutputCamera
               (SetupConcurrentSAD(32), H/128-row bands, pageable buffers) through the C++ mirror

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-extra]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "steroscopic-hardware_b200"))

import numpy as np

W, H, B, D = 1920, 1080, 9, 128
FRAMES = 256                      # frame pairs per step per GPU: 1.06 GB of input >> the 126 MB L2
OPS_PER_EVAL = 6                  # 1 abs-diff + 2 + 2 running-sum add/sub + 1 compare-select
METRIC = "disparity_evals_per_sec_1080p_D128_B9"
UNIT = "Mpix*D/s"


def stream_frame(seed, h=H, w=W, ramp=56):
    """cfg3 generator, SURVEY.md §8(d): textured frame, disparity ramp 8..8+ramp, +-2 noise on R."""
    rng = np.random.default_rng(seed)
    T = rng.integers(0, 256, (h, w + 128 + 8 + ramp + 2), dtype=np.uint8).astype(np.uint16)
    T = ((T[:, :-2] + T[:, 1:-1] + T[:, 2:]) // 3).astype(np.uint8)
    delta = 8 + (ramp * np.arange(h)) // h
    L = np.ascontiguousarray(T[:, 128:128 + w])
    R = np.take_along_axis(T, np.arange(w)[None, :] + 128 + delta[:, None], axis=1)
    R = np.clip(R.astype(np.int16) + rng.integers(-2, 3, R.shape), 0, 255).astype(np.uint8)
    return L, R


def mpixd(frames, seconds):
    return W * H * D * frames / seconds / 1e6


GOLDEN = os.path.join(ROOT, "tests", "golden")
UI_BLOCKS = list(range(3, 32, 2)) + [16]          # cmd/components/control.templ:25-27 + the start-up default (params.go:13-18)
UI_DISPARITIES = list(range(16, 257, 16))         # control.templ:71-73
CFG5_PAIRS = 1024                                 # BASELINE.json configs[4]


def p_int_peak():
    try:
        ip = json.load(open(os.path.join(ROOT, "profiles", "r01_int_peaks.json")))
        return (ip["iadd3"]["lane_ops_per_clk_per_sm"] * ip["sms"] * ip["clock_rate_khz"] * 1e3 / 1e12,
                "profiles/r01_int_peaks.json: measured 63.9 IADD3 lane-ops/clk/SM x 148 SMs x 1.965 GHz")
    except Exception:
        return 148 * 64 * 1.965e9 / 1e12, "theoretical 148 SM x 64 lanes x 1.965 GHz"


def frac_of_roofline(w, h, d, us_per_frame, n_gpus=1):
    return OPS_PER_EVAL * w * h * (d + 1) / (us_per_frame * 1e-6) / 1e12 / (p_int_peak()[0] * n_gpus)


def time_device_batch(torch, ctx, dL, dR, dO, w, h, b, d, fb, reps, warm=1):
    """CUDA-event time of `reps` passes over the resident frame set (launches of fb frames), microseconds per frame."""
    st = torch.cuda.current_stream()
    n = dL.shape[0]

    def one_pass():
        for k in range(0, n, fb):
            m = min(fb, n - k)
            ctx.compute_device_batch(m, dL[k].data_ptr(), w, w * h, dR[k].data_ptr(), w, w * h, w, h, b, d,
                                     dO[k].data_ptr(), w, w * h, cuda_stream=st.cuda_stream)
    for _ in range(warm):
        one_pass()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(st)
    for _ in range(reps):
        one_pass()
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * n)


def load_gray(name):
    from PIL import Image
    return np.array(Image.open(os.path.join(GOLDEN, name)), np.uint8)


def sha(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def extra_single_gpu_configs(torch, despair, O, device_index):
    """cfg1, cfg2 and cfg4 on ONE GPU, device-resident, inputs larger than L2 per pass; parity against the committed golden
    fixtures (SURVEY.md §8(c) SHA pins) and the oracle."""
    out = {}
    man = json.load(open(os.path.join(GOLDEN, "manifest.json")))
    ctx = despair.Context([device_index], 3840, 2160, 1)
    try:
        # cfg1: the four testdata pairs (Go-exact gray), block 9, max disparity 64; 1280 frames = 786 MB of input per pass
        tags = list(man["pairs"].keys())
        Ls = [load_gray(f"L_{t}_gray.png") for t in tags]; Rs = [load_gray(f"R_{t}_gray.png") for t in tags]
        n = 1280
        dL = torch.stack([torch.from_numpy(Ls[k % len(tags)]) for k in range(n)]).cuda()
        dR = torch.stack([torch.from_numpy(Rs[k % len(tags)]) for k in range(n)]).cuda()
        dO = torch.zeros_like(dL)
        us = time_device_batch(torch, ctx, dL, dR, dO, 640, 480, 9, 64, 64, 3)
        got = dO[:len(tags)].cpu().numpy()
        ok = all(sha(got[i]) == man["pairs"][t]["b9_d64_sha256"] for i, t in enumerate(tags))
        out["cfg1"] = {"workload": "testdata L/R_00001,00002,00335,01000 (640x480), block 9, max disparity 64", "us_per_frame": us,
                       "frames_per_sec": 1e6 / us, "frac": frac_of_roofline(640, 480, 64, us), "frames_per_launch": 64,
                       "parity": bool(ok), "parity_against": "SHA-256 pins of SURVEY.md §8(c) (tests/golden/manifest.json), all four pairs",
                       "plan": despair.plan_describe(640, 480, 9, 64, frames=64)}
        del dL, dR, dO
        # cfg2: im0/im1 at 1920x1080, block 15, max disparity 256: intended luma, and the all-zero images LoadPNG really yields
        L = load_gray("im0_intended_gray.png"); R = load_gray("im1_intended_gray.png")
        n = 64
        dL = torch.from_numpy(L).cuda().unsqueeze(0).repeat(n, 1, 1).contiguous(); dR = torch.from_numpy(R).cuda().unsqueeze(0).repeat(n, 1, 1).contiguous()
        dO = torch.zeros_like(dL)
        us = time_device_batch(torch, ctx, dL, dR, dO, 1920, 1080, 15, 256, 16, 2)
        ok = sha(dO[n - 1].cpu().numpy()) == man["survey_pins"]["im0_im1_intended_b15_d256"]
        dL.zero_(); dR.zero_()
        us0 = time_device_batch(torch, ctx, dL, dR, dO, 1920, 1080, 15, 256, 16, 2)
        ok0 = sha(dO[0].cpu().numpy()) == man["survey_pins"]["zeros_1920x1080"]
        out["cfg2"] = {"workload": "im0/im1 1920x1080, block 15, max disparity 256", "us_per_frame": us, "frames_per_sec": 1e6 / us,
                       "frac": frac_of_roofline(1920, 1080, 256, us), "frames_per_launch": 16, "parity": bool(ok),
                       "parity_against": "SHA-256 pin of the intended-luma map (SURVEY.md §8(c))",
                       "faithful_zero_images": {"us_per_frame": us0, "parity": bool(ok0),
                                                "note": "LoadPNG as written turns 8-bit RGB files into all-zero images (gray.go:35-37): the map is all zero"},
                       "plan": despair.plan_describe(1920, 1080, 15, 256, frames=16)}
        del dL, dR, dO
        # cfg4 on one GPU: 3840x2160, block 31, max disparity 256; 16 frames = 265 MB of input per pass
        L, R = stream_frame(4321, 2160, 3840, 240)
        n = 16
        dL = torch.from_numpy(L).cuda().unsqueeze(0).repeat(n, 1, 1).contiguous(); dR = torch.from_numpy(R).cuda().unsqueeze(0).repeat(n, 1, 1).contiguous()
        dO = torch.zeros_like(dL)
        us = time_device_batch(torch, ctx, dL, dR, dO, 3840, 2160, 31, 256, 8, 2)
        rows = dO[n - 1, 1000:1016].cpu().numpy()
        ok = np.array_equal(rows, O.frame_box(L, R, 31, 256, 1000, 1016))
        out["cfg4_one_gpu"] = {"workload": "synthetic 3840x2160, block 31, max disparity 256, device-resident", "us_per_frame": us,
                               "frames_per_sec": 1e6 / us, "frac": frac_of_roofline(3840, 2160, 256, us), "frames_per_launch": 8,
                               "parity": bool(ok), "parity_against": "oracle rows 1000..1015",
                               "plan": despair.plan_describe(3840, 2160, 31, 256, frames=8)}
        del dL, dR, dO
        torch.cuda.empty_cache()
    finally:
        ctx.close()
    return out


def extra_cfg4_sharded(despair, O, n_devices):
    """cfg4 as BASELINE.json names it: ONE 3840x2160 pair, block 31, max disparity 256, row bands + 15-row halo over n_devices GPUs
    of one process (sadgpu_compute_sharded), host buffers in, host-side gather out.  Latency of one frame, pinned and pageable."""
    L, R = stream_frame(4321, 2160, 3840, 240)
    ctx = despair.Context(list(range(n_devices)), 3840, 2160, max(n_devices, 4 if n_devices == 1 else n_devices))
    try:
        pl, pr = ctx.host_pair(2160, 3840); pl[:] = L; pr[:] = R
        po = ctx.host_array((2160, 3840))
        res = {}
        for name, (a, b, o) in (("pinned", (pl, pr, po)), ("pageable", (L, R, np.zeros((2160, 3840), np.uint8)))):
            for _ in range(2):
                ctx.compute_sharded(a, b, 31, 256, out=o)
            t0 = time.perf_counter(); reps = 8
            for _ in range(reps):
                ctx.compute_sharded(a, b, 31, 256, out=o)
            res[name] = (time.perf_counter() - t0) / reps
            if name == "pinned":
                ok = np.array_equal(o[1000:1016], O.frame_box(L, R, 31, 256, 1000, 1016)) and \
                    np.array_equal(o[2150:2160], O.frame_box(L, R, 31, 256, 2150, 2160))
        return {"workload": "ONE synthetic 3840x2160 pair per call, block 31, max disparity 256, row bands + 15-row halo, host buffers",
                "n_gpus": n_devices, "ms_per_frame_pinned": res["pinned"] * 1e3, "ms_per_frame_pageable": res["pageable"] * 1e3,
                "frames_per_sec": 1 / res["pinned"], "Mpix_D_per_s": 3840 * 2160 * 256 / res["pinned"] / 1e6, "parity": bool(ok),
                "parity_against": "oracle rows 1000..1015 and 2150..2159 (band borders included)", "collective": "none (host-side gather)"}
    finally:
        ctx.close()


def extra_cfg5_sweep(torch, dist, despair, O, ctx, rank, world, gen):
    """cfg5: 1024 frame pairs of the cfg3 generator, frame-sharded over the ranks (frame k -> rank k mod world), the full WebUI
    grid.  Every rank times its own share per point with CUDA events; the job time of a point is the max over ranks."""
    from despair.sharding import frames_for_rank
    mine = list(frames_for_rank(CFG5_PAIRS, rank, world))       # frame k of the stream -> rank k mod world
    per_rank = len(mine)
    nuniq = len(gen)
    dL = torch.empty((per_rank, H, W), dtype=torch.uint8, device="cuda"); dR = torch.empty_like(dL)
    for i, k in enumerate(mine):
        dL[i].copy_(torch.from_numpy(gen[k % nuniq][0])); dR[i].copy_(torch.from_numpy(gen[k % nuniq][1]))
    dO = torch.zeros_like(dL)
    points = [(b, d) for b in UI_BLOCKS for d in UI_DISPARITIES]
    times = torch.zeros(len(points), dtype=torch.float64, device="cuda")
    parity = []
    for i, (b, d) in enumerate(points):
        us = time_device_batch(torch, ctx, dL, dR, dO, W, H, b, d, 16, 1, warm=0 if per_rank >= 64 else 1)
        times[i] = us * per_rank                              # microseconds this rank needed for its share
        if rank == 0:
            kl = mine[-1] % nuniq
            parity.append(bool(np.array_equal(dO[per_rank - 1, 520:528].cpu().numpy(), O.frame_box(gen[kl][0], gen[kl][1], b, d, 520, 528))))
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    del dL, dR, dO
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    t = times.cpu().numpy()
    pts = []
    for i, (b, d) in enumerate(points):
        us_job = float(t[i])
        pts.append({"B": b, "D": d, "frames_per_sec": CFG5_PAIRS / (us_job * 1e-6), "us_per_frame_per_gpu": us_job / (CFG5_PAIRS / world),
                    "frac": frac_of_roofline(W, H, d, us_job / (CFG5_PAIRS / world)), "parity": parity[i],
                    "variant": despair.plan_describe(W, H, b, d, frames=16)["variant"]})
    fr = [p["frac"] for p in pts]
    return {"workload": f"{CFG5_PAIRS} synthetic 1080p pairs (cfg3 generator) frame-sharded over {world} GPU(s), {len(points)} points: "
                        f"block {{3..31 odd, 16}} x max disparity {{16..256 step 16}}", "frames_per_launch": 16,
            "points": pts, "all_parity": bool(all(parity)), "parity_against": "oracle rows 520..527 of the last frame of rank 0's share",
            "frac_min": min(fr), "frac_median": float(np.median(fr)), "frac_max": max(fr),
            "sweep_seconds": float(t.sum() * 1e-6)}


def extra_output_camera_path(despair, O):
    """The UNCHANGED reference call pattern (pkg/camera/output.go:129-210 minus the PNG files) through the C++ mirror of pkg/despair:
    SetupConcurrentSAD(32), every frame fresh pageable image objects, H/128-row bands, AssembleDisparityMap; against ONE synchronous
    sadgpu_compute call of the same pageable frame."""
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, "steroscopic-hardware_b200", "libdespair_host.so"))
    u8p = ctypes.c_void_p
    lib.despair_host_output_camera_loop.argtypes = [u8p, u8p] + [ctypes.c_int] * 9 + [u8p, ctypes.POINTER(ctypes.c_double)]
    rng = np.random.default_rng(7)
    legs = []
    ctx = despair.Context([0], 1920, 1080, 1)
    try:
        for (w, h) in ((640, 480), (1920, 1080)):
            b, d, n, warm, iters = 16, 64, 4, 5, 40           # the reference's start-up parameters (params.go:13-18)
            base = rng.integers(0, 256, (n, h, w + 64), dtype=np.uint8)
            L = np.ascontiguousarray(base[:, :, 64:]); R = np.ascontiguousarray(np.roll(base, -19, 2)[:, :, 64:])
            o = np.zeros((h, w), np.uint8)
            for _ in range(5):
                ctx.compute(L[0], R[0], b, d, out=o)
            t0 = time.perf_counter(); reps = 40
            for k in range(reps):
                ctx.compute(L[k % n], R[k % n], b, d, out=o)
            one_call = (time.perf_counter() - t0) / reps * 1e6
            us = ctypes.c_double(); last = np.zeros((h, w), np.uint8)
            rc = lib.despair_host_output_camera_loop(L.ctypes.data, R.ctypes.data, n, w, h, b, d, 32, warm, iters, 0, last.ctypes.data, ctypes.byref(us))
            k = (warm + iters - 1) % n
            legs.append({"w": w, "h": h, "block_size": b, "max_disparity": d, "workers": 32, "chunks_per_frame": -(-h // max(1, h // 128)),
                         "us_per_frame": us.value, "frames_per_sec": 1e6 / us.value if us.value else None,
                         "one_sadgpu_compute_call_pageable_us": one_call, "ratio_to_one_call": us.value / one_call,
                         "parity": bool(rc == 0 and np.array_equal(last, O.frame_box(L[k], R[k], b, d)))})
        lib.despair_host_shutdown()
    finally:
        ctx.close()
    return {"how": "despair::SetupConcurrentSAD(32) + row bands of H/128 rows + AssembleDisparityMap (C++ mirror, host/despair.cpp); each chunk is one "
                   "sadgpu_compute_region call, the chunks of a frame share one GPU pass; buffers are pageable std::vector planes",
            "legs": legs}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md).  Started before the
    warm-up (nvidia-smi needs ~1 s to come up), sampled every 20 ms; only samples whose timestamp falls inside
    the timed region are kept."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def wait_ready(self, timeout=5.0):
        t0 = time.time()
        while not self.rows and time.time() - t0 < timeout:
            time.sleep(0.02)

    def stop(self, t_start, t_end):
        if self.proc:
            self.proc.kill()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                pass
        sm, mx, pw, reasons = [], [], [], set()
        for ts, r in self.rows:
            if ts < t_start or ts > t_end + 0.03:
                continue
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(frames_rows, threads, steps, warmup):
    """Times the oracle's literal restatement of pkg/despair (row bands of H/128 rows pulled by a
    thread pool, exactly the production chunking) on a bounded sample: `frames_rows` rows of one
    cfg3 frame per step."""
    from oracle import oracle as O
    O.build()
    L, R = stream_frame(1234)
    y0 = (H - frames_rows) // 2
    for _ in range(warmup):
        O.frame_literal_mt(L, R, B, D, threads=threads, y0=y0, y1=y0 + max(8, frames_rows // 8))
    t0 = time.perf_counter()
    for _ in range(steps):
        out = O.frame_literal_mt(L, R, B, D, threads=threads, y0=y0, y1=y0 + frames_rows)
    dt = time.perf_counter() - t0
    frames = steps * frames_rows / H
    return mpixd(frames, dt), dt / steps, out


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line of the contract goes to the real stdout; everything else a library prints there (NCCL prints its
    version banner on stdout at communicator creation) was diverted to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def pin_to_gpu_numa(local_rank):
    """Run this rank (and first-touch its pinned buffers) on the CPUs local to its GPU: the end-to-end leg is bound by
    host memory and PCIe, and a rank on the far socket pays for every frame twice."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        return sorted(os.sched_getaffinity(0))
    except Exception:
        return None


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                                      # fd 1 -> stderr for the rest of the run; emit() writes the JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the legs for the other BASELINE configs (headline only)")
    ap.add_argument("--e2e-streams", type=int, default=4, help="calls in flight per GPU in the end-to-end leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    workload = {"workload": "cfg3 synthetic 1920x1080 stereo stream, block 9, max disparity 128",
                "frames_per_step_per_gpu": args.frames, "sharding": "frames across ranks, no collective",
                "l2_policy": "inputs larger than L2 (2*frames*W*H bytes streamed per step)"}

    if args.impl == "reference":
        # The reference's own CPU implementation of the path (oracle port: Go toolchain absent), all host threads.
        if rank != 0:
            return
        # bounded sample: about a quarter frame per step, a whole number of rounds of the thread pool over the production bands
        # (H/128 = 8 rows each, output.go:172) so that no thread idles in the last round
        band = max(1, H // 128)
        rows = min(H, cores * band * max(1, round(270 / (cores * band))))
        v, sec, _ = cpu_reference_run(rows, cores, max(1, args.steps), min(args.warmup, 1))
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload,
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": f"{rows} rows of one cfg3 frame per step, literal O(B^2 D) algorithm, "
                                           f"row bands of H/128 rows on {cores} pthreads"},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    import torch
    import torch.distributed as dist
    import despair
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the SAD path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    affinity = pin_to_gpu_numa(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_streams = max(1, args.e2e_streams)
    ctx = despair.Context([local_rank], W, H, n_streams)

    # ---- synthetic inputs: 8 distinct generated frames tiled to FRAMES (generation is slow on the host) ----
    nuniq = 8
    gen = [stream_frame(1234 + rank * 1024 + k) for k in range(nuniq)]
    F = args.frames
    dL = torch.empty((F, H, W), dtype=torch.uint8, device="cuda")
    dR = torch.empty((F, H, W), dtype=torch.uint8, device="cuda")
    for k in range(F):
        dL[k].copy_(torch.from_numpy(gen[k % nuniq][0])); dR[k].copy_(torch.from_numpy(gen[k % nuniq][1]))
    dO = torch.zeros((F, H, W), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()
    st = stream.cuda_stream

    FB = min(F, 16)                   # frames per launch (sadgpu_compute_device_batch)

    def step_device():
        for k in range(0, F, FB):
            n = min(FB, F - k)
            ctx.compute_device_batch(n, dL[k].data_ptr(), W, W * H, dR[k].data_ptr(), W, W * H, W, H, B, D,
                                     dO[k].data_ptr(), W, W * H, cuda_stream=st)


    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank); sampler.start()
    for _ in range(max(3, args.warmup)):
        step_device()
    launches_per_batch = ctx.last_launch_count()
    batches_per_step = (F + FB - 1) // FB
    barrier()
    sampler.wait_ready()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    t_start = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    t_end = time.time()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t_start, t_end)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = mpixd(args.steps * F * world, ms_max * 1e-3)

    # ---- parity of the timed work (one frame vs oracle rows) ----
    parity = None
    if rank == 0:
        from oracle import oracle as O
        O.build()
        exp = O.frame_box(gen[0][0], gen[0][1], B, D, 520, 536)
        parity = bool(np.array_equal(dO[0].cpu().numpy()[520:536], exp))

    # ---- e2e: C ABI with host buffers (pinned pool): the video-stream call sadgpu_submit_batch_into (EB frame pairs per
    #      call: one H2D DMA, one launch, one D2H DMA) pipelined over n_streams; every frame crosses PCIe both ways inside
    #      the timed region.  The single-frame call (sadgpu_submit_into) is timed beside it. ----
    EB = 8
    ctx.reserve_batch(EB)
    nbuf = 2 * n_streams                                   # distinct pinned input batches (all 8 generated frames appear)
    pin = [ctx.host_array((EB, 2, H, W)) for _ in range(nbuf)]
    for b in range(nbuf):
        for f in range(EB):
            pin[b][f, 0] = gen[(b * EB + f) % nuniq][0]; pin[b][f, 1] = gen[(b * EB + f) % nuniq][1]
    outs = [ctx.host_array((EB, H, W)) for _ in range(n_streams)]
    nbatch = F // EB

    # A step is one pass over the F frames; the stream is continuous, so the calls still in flight at the end of a step are
    # waited for at the start of the next one (as a camera pipeline would) and ALL of them before the clock stops.
    tickets = [None] * n_streams

    def drain():
        for s in range(n_streams):
            if tickets[s] is not None:
                ctx.wait(tickets[s]); tickets[s] = None

    def step_e2e():
        for k in range(nbatch):
            s = k % n_streams
            if tickets[s] is not None:
                ctx.wait(tickets[s])
            tickets[s] = ctx.submit_batch(pin[k % nbuf], B, D, outs[s], stream=s)

    def step_e2e_single():
        for k in range(F):
            s = k % n_streams
            if tickets[s] is not None:
                ctx.wait(tickets[s])
            pb = pin[(k // EB) % nbuf]
            tickets[s] = ctx.submit(pb[k % EB, 0], pb[k % EB, 1], B, D, stream=s, out=outs[s][0])

    def timed_host(fn, steps):
        fn(); drain()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        drain()                                             # every result is in host memory before the clock stops
        torch.cuda.synchronize()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    e2e_steps = max(1, min(args.steps, 5))
    dt_single = timed_host(step_e2e_single, e2e_steps)
    dt_e2e = timed_host(step_e2e, e2e_steps)
    e2e_frames = e2e_steps * nbatch * EB * world
    e2e_value = mpixd(e2e_frames, dt_e2e)
    e2e_parity = None
    if rank == 0:
        kb = (nbatch - 1) % nbuf                            # the last batch written to outs[(nbatch-1) % n_streams]
        fl = (kb * EB + EB - 1) % nuniq
        e2e_parity = bool(np.array_equal(outs[(nbatch - 1) % n_streams][EB - 1][520:536],
                                         O.frame_box(gen[fl][0], gen[fl][1], B, D, 520, 536)))

    # ---- e2e from PAGEABLE caller memory: the synchronous drop-in call sadgpu_compute on plain numpy planes (what a Go Pix slice
    #      is), one frame per call, the staging copies inside the call (rank 0's GPU only) ----
    pageable = None
    if rank == 0:
        po = np.zeros((H, W), np.uint8)
        for k in range(3):
            ctx.compute(gen[k][0], gen[k][1], B, D, out=po)
        t0 = time.perf_counter(); reps = 40
        for k in range(reps):
            ctx.compute(gen[k % nuniq][0], gen[k % nuniq][1], B, D, out=po)
        dtp = (time.perf_counter() - t0) / reps
        pageable = {"value": mpixd(1, dtp), "unit": UNIT, "frames_per_sec": 1 / dtp, "us_per_call": dtp * 1e6, "n_gpus": 1,
                    "parity": bool(np.array_equal(po[520:536], O.frame_box(gen[(reps - 1) % nuniq][0], gen[(reps - 1) % nuniq][1], B, D, 520, 536))),
                    "how": "sadgpu_compute, one 1080p pair per call from pageable numpy planes into a pageable map (staging inside the call)"}

    # ---- the other BASELINE configs ----
    def wait_for_rank0(tag, work):
        """Rank 0 runs `work`; the other ranks sleep on the rendezvous store (a CPU wait: an NCCL barrier would park a spinning
        kernel on every GPU and a spinning thread on every rank's core while rank 0 measures)."""
        barrier()
        res = None
        if world == 1:
            return work()
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            try:
                res = work()
            finally:
                store.set(tag, "done")
        else:
            store.wait([tag])
        return res

    extra = None
    if not args.no_extra:
        del dL, dR, dO
        torch.cuda.empty_cache()
        extra = {}
        r0 = wait_for_rank0("extra_single", lambda: {**extra_single_gpu_configs(torch, despair, O, local_rank),
                                                     "output_camera_path": extra_output_camera_path(despair, O)})
        if rank == 0:
            extra.update(r0)
        barrier()
        sweep = extra_cfg5_sweep(torch, dist, despair, O if rank == 0 else None, ctx, rank, world, gen)
        # cfg4 row-band sharded over ALL the GPUs of the job from ONE process (rank 0); the other ranks' GPUs must be idle
        bands = wait_for_rank0("extra_cfg4", lambda: extra_cfg4_sharded(despair, O, world))
        if rank == 0:
            extra["cfg5_sweep"] = sweep
            extra["cfg4_row_bands"] = bands
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (per launch == per frame), integer ALU ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    p_int, p_src = p_int_peak()
    us_per_frame = ms_max * 1e3 / (args.steps * F)
    us_per_launch = us_per_frame * FB
    evals = W * H * (D + 1) * FB                   # evaluations one launch processes
    achieved = OPS_PER_EVAL * evals / (us_per_launch * 1e-6) / 1e12
    alg_bytes = 3 * W * H * FB
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "latest_traffic.json")))
        traffic = tj["dram_bytes_per_launch"] * FB / tj["frames_per_launch"]     # ncu capture of one 16-frame launch
    except Exception:
        pass
    roofline = {"bound": "int_alu", "achieved": achieved, "peak": p_int, "unit": "Tiop/s", "frac": achieved / p_int,
                "traffic": traffic, "ops_per_eval": OPS_PER_EVAL, "evals_per_launch": evals,
                "kernel_us_per_launch": us_per_launch, "frames_per_launch": FB, "peak_source": p_src,
                "hbm": {"achieved": alg_bytes / (us_per_launch * 1e-6) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": alg_bytes / (us_per_launch * 1e-6) / 1e9 / hbm_peak, "algorithmic_bytes": alg_bytes,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s"}}

    cpu_baseline = None
    if args.gpus == 1 and not args.no_cpu_baseline:
        band = max(1, H // 128)                        # half a frame, a whole number of rounds of the thread pool (see the reference arm)
        rows = min(H, cores * band * max(1, round(540 / (cores * band))))
        v, sec, out = cpu_reference_run(rows, cores, 1, 0)
        ok = bool(np.array_equal(out, ctx.compute(*stream_frame(1234), B, D)[(H - rows) // 2:(H - rows) // 2 + rows]))
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{rows} rows of one cfg3 frame ({sec:.2f} s), literal pkg/despair algorithm "
                                  f"(oracle, Go toolchain unavailable), H/128-row bands on {cores} pthreads",
                        "matches_gpu": ok}

    # the copy-only ceiling of this class of box at N GPUs (tools/pcie_ngpu.py: the same transfers, no kernels)
    ceiling = None
    try:
        pj = json.load(open(os.path.join(ROOT, "profiles", "r02_pcie_ngpu.json")))
        for r in pj["runs"]:
            if r["n_gpus"] == world and r["direction"] == "both" and r["driver"] == "one process per GPU":
                ceiling = {"copy_only_frame_pairs_per_sec": r["frame_pairs_per_s"], "total_GBps": r["total_GBps"],
                           "e2e_fraction_of_ceiling": (e2e_frames / dt_e2e) / r["frame_pairs_per_s"],
                           "source": "profiles/r02_pcie_ngpu.json (8-pair H2D + 8-map D2H copies only, 4 in flight per GPU, one process per GPU)"}
    except Exception:
        pass

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload,
            "frames_per_sec": args.steps * F * world / (ms_max * 1e-3), "us_per_frame_per_gpu": us_per_frame,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * W * H * nbatch * EB * world,
                    "d2h_bytes_per_step": W * H * nbatch * EB * world,
                    "frames_per_sec": e2e_frames / dt_e2e, "parity": e2e_parity,
                    "how": f"sadgpu_submit_batch_into/sadgpu_wait: {EB} frame pairs per call (one H2D DMA, one launch, one D2H DMA), "
                           f"pinned host buffers both ways, {n_streams} streams in flight",
                    "single_frame_calls": {"value": mpixd(e2e_steps * F * world, dt_single), "unit": UNIT,
                                           "frames_per_sec": e2e_steps * F * world / dt_single,
                                           "how": "sadgpu_submit_into/sadgpu_wait, one frame pair per call"},
                    "pageable": pageable, "pcie_ceiling": ceiling},
            "gpu_launches": args.steps * batches_per_step * launches_per_batch, "frames_per_launch": FB,
            "host": {"cores": cores, "rank0_cpu_affinity": (f"{affinity[0]}-{affinity[-1]} ({len(affinity)} cpus, GPU-local)" if affinity else "unpinned")},
            "parity": parity, "plan": despair.plan_describe(W, H, B, D, frames=FB), "clocks": clocks,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "extra": extra}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
