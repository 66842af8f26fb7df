"""The UNCHANGED reference call pattern (pkg/camera/output.go:129-210: SetupConcurrentSAD(32), H/128-row bands, AssembleDisparityMap)
through the C++ mirror, pageable buffers, fresh image objects per frame, against ONE synchronous sadgpu_compute call of the same frame.
    python tools/camera_path.py  -> gpurun_out/camera_path.json"""
import ctypes, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "steroscopic-hardware_b200"))
import numpy as np, despair

H_LIB = ctypes.CDLL(os.path.join(ROOT, "steroscopic-hardware_b200", "libdespair_host.so"))
u8p = ctypes.c_void_p
H_LIB.despair_host_output_camera_loop.argtypes = [u8p, u8p] + [ctypes.c_int] * 9 + [u8p, ctypes.POINTER(ctypes.c_double)]


def main():
    rng = np.random.default_rng(0)
    out = []
    ctx = despair.Context([0], 3840, 2160, 1)
    for (W, H, B, D) in ((640, 480, 16, 64), (640, 480, 9, 64), (1920, 1080, 16, 64), (1920, 1080, 9, 128)):
        n = 4
        base = rng.integers(0, 256, (n, H, W + 64), dtype=np.uint8)
        L = np.ascontiguousarray(base[:, :, 64:]); R = np.ascontiguousarray(np.roll(base, -19, 2)[:, :, 64:])
        o = np.zeros((H, W), np.uint8)
        for _ in range(5): ctx.compute(L[0], R[0], B, D, out=o)
        t0 = time.perf_counter(); reps = 50
        for k in range(reps): ctx.compute(L[k % n], R[k % n], B, D, out=o)
        one_call = (time.perf_counter() - t0) / reps * 1e6
        rec = {"W": W, "H": H, "B": B, "D": D, "sadgpu_compute_pageable_us": round(one_call, 1)}
        for workers in (32, 8):
            for reuse in (0, 1):
                us = ctypes.c_double(); last = np.zeros((H, W), np.uint8)
                rc = H_LIB.despair_host_output_camera_loop(L.ctypes.data, R.ctypes.data, n, W, H, B, D, workers, 5, 40, reuse, last.ctypes.data, ctypes.byref(us))
                assert rc == 0
                same = bool(np.array_equal(last, ctx.compute(L[(5 + 40 - 1) % n], R[(5 + 40 - 1) % n], B, D)))
                rec[f"output_camera_path_us_workers{workers}_{'reused' if reuse else 'fresh'}_objects"] = round(us.value, 1)
                rec[f"matches_single_call_workers{workers}_{reuse}"] = same
        rec["chunks_per_frame"] = -(-H // max(1, H // 128))
        out.append(rec); print(rec, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "camera_path.json"), "w"), indent=1)


main()
