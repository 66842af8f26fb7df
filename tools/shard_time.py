"""cfg4 (3840x2160, B=31, D=256): one frame split into row bands with a block_size/2 halo over N devices of ONE process
(sadgpu_compute_sharded: host buffers in, host-side gather out) for N = 1, 2, 4, 8 -> gpurun_out/shard_times.json"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "steroscopic-hardware_b200"))
import numpy as np, torch, despair
W, H, B, D = 3840, 2160, 31, 256
rng = np.random.default_rng(4321)
T = rng.integers(0, 256, (H, W + 512), dtype=np.uint8)
L = np.ascontiguousarray(T[:, 256:256 + W]); R = np.ascontiguousarray(np.roll(T, -37, 1)[:, 256:256 + W])
ndev = torch.cuda.device_count()
ref_rows = None          # rows of the one-band run: every sharded run must reproduce them (the oracle comparison lives in tests/)
out = []
for n in (1, 2, 4, 8):
    if n > ndev: break
    ctx = despair.Context(list(range(n)), W, H, n)
    pl, pr = ctx.host_pair(H, W); pl[:] = L; pr[:] = R
    po = ctx.host_array((H, W))
    for _ in range(2): ctx.compute_sharded(pl, pr, B, D, out=po)
    t0 = time.perf_counter(); reps = 10
    for _ in range(reps): ctx.compute_sharded(pl, pr, B, D, out=po)
    dt = (time.perf_counter() - t0) / reps
    if ref_rows is None: ref_rows = po[1000:1100].copy()
    ok = bool(np.array_equal(po[1000:1100], ref_rows))
    r = {"n_gpus": n, "ms_per_frame": round(dt * 1e3, 3), "frames_per_sec": round(1 / dt, 1), "Mpix_D_per_s": round(W * H * D / dt / 1e6, 1), "rows_match_single_gpu_run": ok}
    out.append(r); print(r, flush=True)
    ctx.close()
json.dump({"config": "cfg4 3840x2160 B=31 D=256, one frame per call, pinned host buffers, row bands + 15-row halo, no collective", "runs": out},
          open(os.path.join(ROOT, "gpurun_out", "shard_times.json"), "w"), indent=1)
