"""PCIe probe: pinned H2D / D2H bandwidth by transfer size, concurrency and host allocation flags (developer tool)."""
import ctypes, subprocess, time
import torch

rt = ctypes.CDLL("libcudart.so.12")
rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]

def host_alloc(n, flags):
    p = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(p), n, flags) == 0
    return p

def timeit(fn, reps):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps

print(subprocess.run("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current,pcie.link.width.max --format=csv",
                     shell=True, capture_output=True, text=True).stdout)
torch.cuda.init(); torch.zeros(1, device="cuda")
MB = 1 << 20
streams = [torch.cuda.Stream() for _ in range(4)]
dev = torch.empty(512 * MB, dtype=torch.uint8, device="cuda")
for name, flags in (("default", 0), ("portable", 1), ("write-combined", 4)):
    hp = host_alloc(512 * MB, flags)
    ctypes.memset(hp, 1, 512 * MB)
    for size in (2 * MB, 4 * MB, 8 * MB, 32 * MB, 128 * MB):
        reps = max(10, min(400, (2048 * MB) // size))
        def h2d(k=1):
            for i in range(k):
                rt.cudaMemcpyAsync(dev.data_ptr() + i * size, hp.value + i * size, size, 1, streams[i].cuda_stream)
        def d2h(k=1):
            for i in range(k):
                rt.cudaMemcpyAsync(hp.value + (8 + i) * size % (256 * MB), dev.data_ptr() + 256 * MB + i * size, size, 2, streams[2 + i].cuda_stream)
        r = {}
        r["h2d x1"] = size / timeit(lambda: h2d(1), reps) / 1e9
        r["h2d x2"] = 2 * size / timeit(lambda: h2d(2), reps) / 1e9
        r["d2h x1"] = size / timeit(lambda: d2h(1), reps) / 1e9
        dt = timeit(lambda: (h2d(1), d2h(1)), reps)
        r["h2d+d2h (GB/s each way)"] = size / dt / 1e9
        dt = timeit(lambda: (h2d(2), d2h(1)), reps)
        r["2h2d+d2h (h2d GB/s)"] = 2 * size / dt / 1e9
        print(f"{name:15s} {size // MB:4d} MB  " + "  ".join(f"{k}: {v:5.1f}" for k, v in r.items()), flush=True)
