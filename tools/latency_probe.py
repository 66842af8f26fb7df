"""Latency of the synchronous drop-in call sadgpu_compute (what RunSad / one OutputCamera frame costs): pageable vs pinned
buffers, the reference's frame sizes and start-up parameters."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "steroscopic-hardware_b200"))
import numpy as np, despair
rng = np.random.default_rng(0)
ctx = despair.Context([0], 3840, 2160, 4)
for (W, H, B, D) in ((640, 480, 9, 64), (640, 480, 16, 64), (1920, 1080, 9, 128), (1920, 1080, 16, 64), (3840, 2160, 31, 256)):
    L = rng.integers(0, 256, (H, W), dtype=np.uint8); R = np.roll(L, -11, 1)
    out = np.zeros((H, W), np.uint8)
    pl, pr = ctx.host_pair(H, W); pl[:] = L; pr[:] = R; po = ctx.host_array((H, W))
    res = {}
    for name, (a, b, o) in (("pageable", (L, R, out)), ("pinned", (pl, pr, po))):
        for _ in range(5): ctx.compute(a, b, B, D, out=o)
        t0 = time.perf_counter(); n = 50
        for _ in range(n): ctx.compute(a, b, B, D, out=o)
        res[name] = (time.perf_counter() - t0) / n * 1e6
    for name, (a, b, o) in (("pageable", (L, R, out)), ("pinned", (pl, pr, po))):
        for _ in range(5): ctx.compute_sharded(a, b, B, D, out=o)
        t0 = time.perf_counter(); n = 50
        for _ in range(n): ctx.compute_sharded(a, b, B, D, out=o)
        res["banded " + name] = (time.perf_counter() - t0) / n * 1e6
    print(f"{W}x{H} B={B} D={D}: sadgpu_compute {res['pageable']:.0f} us pageable, {res['pinned']:.0f} us pinned; "
          f"sadgpu_compute_sharded (4 bands on one device) {res['banded pageable']:.0f} / {res['banded pinned']:.0f} us  "
          f"({despair.plan_describe(W, H, B, D)['variant']})", flush=True)
