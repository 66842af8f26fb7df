"""numpy statement of the optional post-processing hooks.  The reference has no such stage, so there is nothing to
restate: this file DEFINES the semantics the CUDA hooks are tested against.  TEST INFRASTRUCTURE ONLY."""
import numpy as np


def median3(m: np.ndarray) -> np.ndarray:
    """3x3 median, window clamped to the image (edge pixels replicate)."""
    p = np.pad(m, 1, mode="edge")
    h, w = m.shape
    stack = np.stack([p[dy:dy + h, dx:dx + w] for dy in range(3) for dx in range(3)])
    return np.sort(stack, axis=0)[4].astype(np.uint8)


def lr_check(left_map: np.ndarray, right_map: np.ndarray, max_disparity: int, tolerance: int, invalid: int) -> np.ndarray:
    """left_map / right_map hold v = d*255/D.  A left-map pixel survives iff the right-referenced map at x - d decodes to within
    `tolerance` of d (d = round(v*D/255)); pixels whose match leaves the image are invalid."""
    h, w = left_map.shape
    d = (left_map.astype(np.int64) * max_disparity + 127) // 255
    x = np.arange(w)[None, :] - d
    ok = x >= 0
    dr = (np.take_along_axis(right_map, np.clip(x, 0, w - 1), axis=1).astype(np.int64) * max_disparity + 127) // 255
    ok &= np.abs(dr - d) <= tolerance
    return np.where(ok, left_map, invalid).astype(np.uint8)


def right_referenced(frame_fn, left: np.ndarray, right: np.ndarray, block_size: int, max_disparity: int) -> np.ndarray:
    """The right-referenced map = the path on the mirrored pair with the roles swapped, mirrored back."""
    return np.ascontiguousarray(frame_fn(np.ascontiguousarray(right[:, ::-1]), np.ascontiguousarray(left[:, ::-1]),
                                         block_size, max_disparity)[:, ::-1])
