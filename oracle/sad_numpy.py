"""numpy twin of the oracle (independent of sad_oracle.c) — TEST INFRASTRUCTURE ONLY.

box_numpy   : closed form of pkg/despair/sad.go:55-95 + :205-244 using
              an integral image per disparity.
literal_py  : pure-Python transliteration of the two Go functions, for tiny cases.
"""
import numpy as np


def box_numpy(left: np.ndarray, right: np.ndarray, block_size: int, max_disparity: int) -> np.ndarray:
    L = left.astype(np.int64)
    R = right.astype(np.int64)
    H, W = L.shape
    h = block_size // 2
    best = np.full((H, W), np.iinfo(np.int64).max, np.int64)
    bestd = np.zeros((H, W), np.int64)
    xs = np.arange(W)
    for d in range(max_disparity + 1):
        ad = np.zeros((H, W), np.int64)
        if d < W:
            ad[:, d:] = np.abs(L[:, d:] - R[:, :W - d])
        ii = np.zeros((H + 1, W + 1), np.int64)
        ii[1:, 1:] = ad.cumsum(0).cumsum(1)
        y0 = np.clip(np.arange(H) - h, 0, H)[:, None]
        y1 = np.clip(np.arange(H) + h + 1, 0, H)[:, None]
        x0 = np.clip(xs - h, 0, W)[None, :]
        x1 = np.clip(xs + h + 1, 0, W)[None, :]
        S = ii[y1, x1] - ii[y0, x1] - ii[y1, x0] + ii[y0, x0]
        valid = (xs >= h + d)[None, :]          # candidates: d <= X - h  (and X >= h)
        upd = valid & (S < best)
        best[upd] = S[upd]
        bestd[upd] = d
    return ((bestd * 255) // max_disparity).astype(np.uint8)


def _sad_py(Lp, Rp, W, H, lx, ly, rx, ry, B):
    hs = B // 2
    lminy = max(ly - hs, 0); lmaxy = min(ly + hs + 1, H)
    lminx = max(lx - hs, 0); lmaxx = min(lx + hs + 1, W)
    rminy = max(ry - hs, 0); rminx = max(rx - hs, 0)
    sad = 0
    for y in range(lminy, lmaxy):
        if rminy + (y - lminy) >= H:
            break
        for x in range(lminx, lmaxx):
            if rminx + (x - lminx) >= W:
                break
            sad += abs(int(Lp[y][x]) - int(Rp[rminy + (y - lminy)][rminx + (x - lminx)]))
    return sad


def literal_py(left: np.ndarray, right: np.ndarray, block_size: int, max_disparity: int) -> np.ndarray:
    H, W = left.shape
    Lp = left.tolist(); Rp = right.tolist()
    out = np.zeros((H, W), np.uint8)
    for y in range(H):
        for x in range(W):
            m = 2 ** 31 - 1; b = 0
            for d in range(max_disparity + 1):
                if x - d < 0:
                    continue
                s = _sad_py(Lp, Rp, W, H, x, y, x - d, y, block_size)
                if s < m:
                    m = s; b = d
                    if s == 0:
                        break
            out[y, x] = (b * 255) // max_disparity
    return out
