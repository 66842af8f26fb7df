"""Copy-only probe of the host <-> N GPU path (no kernels): the transfers of bench.py's end-to-end leg — per call one pinned H2D copy of 8
1080p frame pairs (33.2 MB) and one D2H copy of 8 maps (16.6 MB), 4 calls in flight per GPU — on 1, 2, 4, 8 GPUs at once, driven
(a) by one process per GPU, as `torchrun bench.py` does, and (b) by ONE process for all GPUs.  Answers whether the end-to-end scaling of
the frame-sharded job is bounded by the box's aggregate PCIe / host-memory bandwidth or by the submit / wait loop.
    python tools/pcie_ngpu.py  -> gpurun_out/pcie_ngpu.json"""
import json, os, sys, time
import torch
import torch.multiprocessing as mp

H2D_BYTES = 8 * 2 * 1920 * 1080
D2H_BYTES = 8 * 1920 * 1080
CALLS, INFLIGHT = 64, 4


def copy_loop(dev, direction, barrier=None):
    """`direction`: 'h2d', 'd2h' or 'both'.  Returns seconds for CALLS calls on device `dev`."""
    torch.cuda.set_device(dev)
    streams = [torch.cuda.Stream(dev) for _ in range(INFLIGHT)]
    hin = [torch.empty(H2D_BYTES, dtype=torch.uint8).pin_memory() for _ in range(INFLIGHT)]
    hout = [torch.empty(D2H_BYTES, dtype=torch.uint8).pin_memory() for _ in range(INFLIGHT)]
    din = [torch.empty(H2D_BYTES, dtype=torch.uint8, device=f"cuda:{dev}") for _ in range(INFLIGHT)]
    dout = [torch.zeros(D2H_BYTES, dtype=torch.uint8, device=f"cuda:{dev}") for _ in range(INFLIGHT)]

    def run(n):
        for k in range(n):
            s = k % INFLIGHT
            streams[s].synchronize()                      # the wait of the call that used this stream before
            with torch.cuda.stream(streams[s]):
                if direction != "d2h":
                    din[s].copy_(hin[s], non_blocking=True)
                if direction != "h2d":
                    hout[s].copy_(dout[s], non_blocking=True)
        for s in streams:
            s.synchronize()
    run(8)
    if barrier is not None:
        barrier.wait()
    t0 = time.perf_counter()
    run(CALLS)
    return time.perf_counter() - t0


def _proc(dev, direction, barrier, q):
    q.put((dev, copy_loop(dev, direction, barrier)))


def one_process_all_gpus(n, direction):
    """One host thread drives all n GPUs (enqueues are asynchronous): what a single-process server sees."""
    res = []
    streams, hin, hout, din, dout = {}, {}, {}, {}, {}
    for d in range(n):
        torch.cuda.set_device(d)
        streams[d] = [torch.cuda.Stream(d) for _ in range(INFLIGHT)]
        hin[d] = [torch.empty(H2D_BYTES, dtype=torch.uint8).pin_memory() for _ in range(INFLIGHT)]
        hout[d] = [torch.empty(D2H_BYTES, dtype=torch.uint8).pin_memory() for _ in range(INFLIGHT)]
        din[d] = [torch.empty(H2D_BYTES, dtype=torch.uint8, device=f"cuda:{d}") for _ in range(INFLIGHT)]
        dout[d] = [torch.zeros(D2H_BYTES, dtype=torch.uint8, device=f"cuda:{d}") for _ in range(INFLIGHT)]

    def run(calls):
        for k in range(calls):
            s = k % INFLIGHT
            for d in range(n):
                streams[d][s].synchronize()
                with torch.cuda.stream(streams[d][s]):
                    if direction != "d2h":
                        din[d][s].copy_(hin[d][s], non_blocking=True)
                    if direction != "h2d":
                        hout[d][s].copy_(dout[d][s], non_blocking=True)
        for d in range(n):
            for s in streams[d]:
                s.synchronize()
    run(8)
    t0 = time.perf_counter()
    run(CALLS)
    return time.perf_counter() - t0


def gbs(n, direction, seconds):
    up = H2D_BYTES * CALLS * n / seconds / 1e9 if direction != "d2h" else 0.0
    down = D2H_BYTES * CALLS * n / seconds / 1e9 if direction != "h2d" else 0.0
    return {"h2d_GBps": round(up, 1), "d2h_GBps": round(down, 1), "total_GBps": round(up + down, 1),
            "frame_pairs_per_s": round(8 * CALLS * n / seconds, 0) if direction == "both" else None}


def main():
    ndev = torch.cuda.device_count()
    out = {"h2d_bytes_per_call": H2D_BYTES, "d2h_bytes_per_call": D2H_BYTES, "calls": CALLS, "in_flight_per_gpu": INFLIGHT,
           "host_cpus": os.cpu_count(), "runs": []}
    ctx = mp.get_context("spawn")
    for n in (1, 2, 4, 8):
        if n > ndev:
            break
        for direction in ("h2d", "d2h", "both"):
            barrier = ctx.Barrier(n); q = ctx.Queue()
            ps = [ctx.Process(target=_proc, args=(d, direction, barrier, q)) for d in range(n)]
            [p.start() for p in ps]
            times = [q.get(timeout=300)[1] for _ in ps]
            [p.join() for p in ps]
            r = {"n_gpus": n, "direction": direction, "driver": "one process per GPU", **gbs(n, direction, max(times))}
            out["runs"].append(r); print(r, flush=True)
            r = {"n_gpus": n, "direction": direction, "driver": "one process, all GPUs", **gbs(n, direction, one_process_all_gpus(n, direction))}
            out["runs"].append(r); print(r, flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/pcie_ngpu.json", "w"), indent=1)


if __name__ == "__main__":
    main()
