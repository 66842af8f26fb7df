"""Host-side partitioning of the SAD path over GPUs / ranks.  No collective on the data
path: the per-pixel function only reads rows Y-h..Y+h, so a row band plus an h-row halo is self-contained."""
from typing import List, Tuple


def row_bands(h: int, n: int, block_size: int) -> List[Tuple[int, int, int, int]]:
    """n row bands of an h-row frame: (y0, y1, ys, ye) = output rows [y0,y1), input rows [ys,ye) incl. halo.
    Mirrors sadgpu_compute_sharded and the reference's own band chunking (pkg/camera/output.go:172-187)."""
    half = block_size // 2
    out = []
    for i in range(n):
        y0, y1 = h * i // n, h * (i + 1) // n
        out.append((y0, y1, max(0, y0 - half), min(h, y1 + half)))
    return out


def frames_for_rank(n_frames: int, rank: int, world: int) -> range:
    """Frame k -> rank k mod world (video streams, BASELINE configs[4])."""
    return range(rank, n_frames, world)
