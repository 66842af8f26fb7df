"""ctypes front-end of oracle/sad_oracle.c plus the pure-host pieces of pkg/despair
(chunk planners and AssembleDisparityMap) restated in Python.

TEST INFRASTRUCTURE ONLY — never imported by the product package.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from typing import List, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_u8p = ctypes.POINTER(ctypes.c_uint8)


def build(force: bool = False) -> str:
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "sad_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    if os.path.exists("/root/reference/hardware/sad.c") and (
            force or not os.path.exists(os.path.join(_HERE, "_ref", "hw_sad"))):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        I = ctypes.c_int
        L.oracle_sum_abs_diff.argtypes = [_u8p, I, I, I, _u8p, I, I, I, I, I, I, I, I]
        L.oracle_sum_abs_diff.restype = I
        L.oracle_region_literal.argtypes = [_u8p, I, _u8p, I, I, I, I, I, I, I, I, I, I, _u8p]
        L.oracle_region_literal.restype = None
        L.oracle_frame_box.argtypes = [_u8p, I, _u8p, I, I, I, I, I, I, I, _u8p, I]
        L.oracle_frame_box.restype = I
        L.oracle_frame_literal_mt.argtypes = [_u8p, I, _u8p, I, I, I, I, I, I, I, I, I, _u8p, I]
        L.oracle_frame_literal_mt.restype = I
        _LIB = L
    return _LIB


def _img(a: np.ndarray) -> Tuple[np.ndarray, "ctypes._Pointer", int]:
    a = np.asarray(a)
    if a.dtype != np.uint8 or a.ndim != 2:
        raise ValueError("expected a 2-D uint8 image")
    if a.strides[1] != 1:
        a = np.ascontiguousarray(a)
    return a, a.ctypes.data_as(_u8p), a.strides[0]


def sum_abs_diff(left, right, lx, ly, rx, ry, block_size) -> int:
    """pkg/despair/sad.go:205-244."""
    l, lp, ls = _img(left)
    r, rp, rs = _img(right)
    return lib().oracle_sum_abs_diff(lp, ls, l.shape[1], l.shape[0], rp, rs, r.shape[1], r.shape[0],
                                     lx, ly, rx, ry, block_size)


def region_literal(left, right, region, block_size, max_disparity, early_exit=True) -> np.ndarray:
    """pkg/despair/sad.go:55-95 for region=(x0,y0,x1,y1); returns (Dy,Dx) uint8."""
    l, lp, ls = _img(left)
    r, rp, rs = _img(right)
    x0, y0, x1, y1 = region
    out = np.zeros((y1 - y0, x1 - x0), np.uint8)
    lib().oracle_region_literal(lp, ls, rp, rs, l.shape[1], l.shape[0], x0, y0, x1, y1,
                                block_size, max_disparity, int(early_exit), out.ctypes.data_as(_u8p))
    return out


def frame_box(left, right, block_size, max_disparity, y0=0, y1=None) -> np.ndarray:
    """Closed-form O(D) oracle for rows [y0,y1)."""
    l, lp, ls = _img(left)
    r, rp, rs = _img(right)
    h, w = l.shape
    y1 = h if y1 is None else y1
    out = np.zeros((y1 - y0, w), np.uint8)
    rc = lib().oracle_frame_box(lp, ls, rp, rs, w, h, block_size, max_disparity, y0, y1,
                                out.ctypes.data_as(_u8p), w)
    if rc != 0:
        raise ValueError(f"oracle_frame_box rc={rc}")
    return out


def frame_literal_mt(left, right, block_size, max_disparity, threads=1, y0=0, y1=None,
                     early_exit=True) -> np.ndarray:
    """Literal algorithm, production row-band chunking (output.go:172-187) on pthreads."""
    l, lp, ls = _img(left)
    r, rp, rs = _img(right)
    h, w = l.shape
    y1 = h if y1 is None else y1
    out = np.zeros((y1 - y0, w), np.uint8)
    rc = lib().oracle_frame_literal_mt(lp, ls, rp, rs, w, h, block_size, max_disparity, y0, y1,
                                       threads, int(early_exit), out.ctypes.data_as(_u8p), w)
    if rc != 0:
        raise ValueError(f"oracle_frame_literal_mt rc={rc}")
    return out


# ---- pure host logic of pkg/despair, restated ------------------------------------------

Rect = Tuple[int, int, int, int]  # (x0, y0, x1, y1), Go image.Rect order


def run_sad_chunks(w: int, h: int, num_cpu: int) -> List[Rect]:
    """Tile planner of RunSad, pkg/despair/sad.go:128-153 (Rect.Min == (0,0))."""
    num_workers = num_cpu * 4                                # :128
    num_chunks = num_workers * 4                             # :129
    chunk_width = int(math.sqrt(float((w * h) // num_chunks)))  # :138-142 (int division first)
    hor = max(1, w // chunk_width)                           # :143 (ZeroDivisionError == Go panic)
    ver = max(1, num_chunks // hor)                          # :144
    chunk_width = w // hor                                   # :145
    chunk_height = h // ver                                  # :146
    if chunk_height == 0:
        raise RuntimeError("reference loops forever: chunkHeight == 0 (sad.go:147)")
    out: List[Rect] = []
    y = 0
    while y < h:                                             # :147
        ye = min(y + chunk_height, h)
        x = 0
        while x < w:                                         # :149
            out.append((x, y, min(x + chunk_width, w), ye))
            x += chunk_width
        y += chunk_height
    return out


def output_camera_chunks(w: int, h: int, workers: int = 32) -> List[Rect]:
    """Row bands of OutputCamera.processDepthMap, pkg/camera/output.go:172-187."""
    chunk = max(1, h // (workers * 4))
    return [(0, y, w, min(y + chunk, h)) for y in range(0, h, chunk)]


def assemble_disparity_map(chunks: Sequence[Tuple[np.ndarray, Rect]], w: int, h: int,
                           n_chunks: int, faithful_bug: bool = True) -> np.ndarray:
    """AssembleDisparityMap, pkg/despair/sad.go:172-202, for chunks in ARRIVAL order.

    faithful_bug=True reproduces `i++; if i >= chunks {break}` (:179-184): the chunk that
    arrives n_chunks-th is received and dropped, so its rectangle stays 0.
    """
    out = np.zeros((h, w), np.uint8)
    i = 0
    for data, (x0, y0, x1, y1) in chunks:
        i += 1
        if faithful_bug and i >= n_chunks:
            break
        out[y0:y1, x0:x1] = np.asarray(data, np.uint8).reshape(y1 - y0, x1 - x0)
    return out
