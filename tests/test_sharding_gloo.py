"""N>1 host logic on CPU: the partition helpers the product uses — despair.sharding.row_bands (the band / halo arithmetic of
sadgpu_compute_sharded) and frames_for_rank (bench.py's frame sharding of the cfg5 stream) — driven by two gloo ranks with
bench.py's barrier + max-over-ranks timing pattern.  There is no GPU here, so the per-band COMPUTE is the oracle standing in for
the kernel: this test is evidence for the partition and the gather, not for the kernels — those are covered on real devices by
tests/test_gpu_parity.py::test_row_band_sharding_across_devices and by bench.py's cfg4 / cfg5 legs under torchrun."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import sys
    for p in (ROOT, os.path.join(ROOT, "steroscopic-hardware_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from despair.sharding import row_bands, frames_for_rank
    rng = np.random.default_rng(5)                       # same frame on every rank
    H, W, B, D = 97, 150, 11, 40
    base = rng.integers(0, 256, (H, W + 48), dtype=np.uint8)
    L = np.ascontiguousarray(base[:, 48:]); R = np.ascontiguousarray(np.roll(base, -9, 1)[:, 48:])
    y0, y1, ys, ye = row_bands(H, world, B)[rank]
    band = O.frame_box(L[ys:ye], R[ys:ye], B, D, y0 - ys, y1 - ys)          # the band only sees its rows + halo
    full = torch.zeros((H, W), dtype=torch.uint8)
    full[y0:y1] = torch.from_numpy(band)
    dist.all_reduce(full, op=dist.ReduceOp.SUM)                              # host-side gather of disjoint rows
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)                # bench.py's max-over-ranks timing
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # frame sharding: every frame is owned by exactly one rank
    owned = torch.zeros(10, dtype=torch.int32)
    for k in frames_for_rank(10, rank, world):
        owned[k] += 1
    dist.all_reduce(owned)
    if rank == 0:
        q.put((np.array_equal(full.numpy(), O.frame_box(L, R, B, D)), float(t.item()), owned.tolist()))
    dist.destroy_process_group()


def test_two_rank_row_band_sharding_gloo(oracle):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok, tmax, owned = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert ok and tmax == 2.0 and owned == [1] * 10


def test_row_bands_cover_and_halo():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "steroscopic-hardware_b200"))
    from despair.sharding import row_bands
    for h, n, B in [(2160, 8, 31), (1080, 3, 9), (7, 8, 31), (480, 1, 16)]:
        bands = row_bands(h, n, B)
        assert bands[0][0] == 0 and bands[-1][1] == h
        for (a, b, s, e), nxt in zip(bands, bands[1:] + [None]):
            assert s == max(0, a - B // 2) and e == min(h, b + B // 2)
            if nxt:
                assert b == nxt[0]
