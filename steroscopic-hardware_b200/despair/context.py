"""Thin object wrapper over the C ABI (one sadgpu_ctx).  numpy in / numpy out for the host
entry points, raw device pointers (e.g. torch.Tensor.data_ptr()) for the device-resident one."""
import ctypes
import json

import numpy as np

from . import _native as N


def _u8_2d(a, name):
    a = np.asarray(a)
    if a.dtype != np.uint8 or a.ndim != 2:
        raise ValueError(f"{name}: expected a 2-D uint8 array (an image.Gray Pix plane)")
    if a.strides[1] != 1 or a.strides[0] < a.shape[1]:
        a = np.ascontiguousarray(a)
    return a


def _pair(left, right):
    l = _u8_2d(left, "left"); r = _u8_2d(right, "right")
    if l.shape != r.shape:
        raise N.SadGpuError(-1)     # left.Rect != right.Rect is rejected (SURVEY.md §8 deviations)
    return l, r


def _out_2d(out, shape, name="out"):
    """A caller-supplied destination must be a writable uint8 plane of the frame's shape with unit column stride:
    the C side writes shape[1] bytes per row at out.strides[0] intervals."""
    if out is None:
        return np.zeros(shape, np.uint8)
    if not isinstance(out, np.ndarray) or out.dtype != np.uint8 or out.shape != tuple(shape) or out.strides[1] != 1 \
            or out.strides[0] < shape[1] or not out.flags.writeable:
        raise ValueError(f"{name}: expected a writable uint8 array of shape {tuple(shape)} with contiguous rows")
    return out


class Context:
    def __init__(self, devices=None, max_w=1920, max_h=1080, n_streams=1):
        L = N.lib()
        if devices is None:
            devices = [0]
        arr = (ctypes.c_int * len(devices))(*devices)
        h = ctypes.c_void_p()
        N.check(L.sadgpu_create(arr, len(devices), max_w, max_h, n_streams, ctypes.byref(h)))
        self._h = h
        self._L = L
        self.devices = list(devices)
        self.n_streams = n_streams
        self.max_w, self.max_h = max_w, max_h

    def close(self):
        if getattr(self, "_h", None):
            self._L.sadgpu_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- host entry points -------------------------------------------------------------
    def compute(self, left, right, block_size, max_disparity, y0=0, y1=None, stream=0, out=None):
        l, r = _pair(left, right)
        h, w = l.shape
        y1 = h if y1 is None else y1
        out = _out_2d(out, (h, w))
        N.check(self._L.sadgpu_compute(self._h, stream, l.ctypes.data, l.strides[0], r.ctypes.data, r.strides[0],
                                        w, h, block_size, max_disparity, y0, y1, out.ctypes.data, out.strides[0]))
        return out

    def compute_region(self, left, right, block_size, max_disparity, region, out=None):
        """One InputChunk (sad.go:12-15): region = (x0, y0, x1, y1) in image coordinates (Go image.Rect order); returns the
        region-local (Dy, Dx) block — OutputChunk.DisparityData.  Chunks of the same frame pair share one GPU pass."""
        l, r = _pair(left, right)
        h, w = l.shape
        x0, y0, x1, y1 = region
        out = _out_2d(out, (max(0, y1 - y0), max(0, x1 - x0)))
        N.check(self._L.sadgpu_compute_region(self._h, l.ctypes.data, l.strides[0], r.ctypes.data, r.strides[0], w, h,
                                               block_size, max_disparity, x0, y0, x1, y1, out.ctypes.data,
                                               out.strides[0] if out.size else max(1, x1 - x0)))
        return out

    def region_stats(self):
        a, b, c = ctypes.c_longlong(), ctypes.c_longlong(), ctypes.c_longlong()
        N.check(self._L.sadgpu_region_stats(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return {"calls": a.value, "frames": b.value, "stale": c.value}

    def compute_nrgba(self, left_rgba, right_rgba, block_size, max_disparity, stream=0, out=None):
        """left/right: uint8 [h][w][4] non-premultiplied RGBA (image.NRGBA.Pix); luma on the device, Go-exact."""
        l = np.asarray(left_rgba); r = np.asarray(right_rgba)
        if l.dtype != np.uint8 or l.ndim != 3 or l.shape[2] != 4 or l.shape != r.shape or r.dtype != np.uint8:
            raise ValueError("expected two uint8 arrays [h][w][4] of the same shape")
        if l.strides[2] != 1 or l.strides[1] != 4: l = np.ascontiguousarray(l)
        if r.strides[2] != 1 or r.strides[1] != 4: r = np.ascontiguousarray(r)
        h, w, _ = l.shape
        out = _out_2d(out, (h, w))
        N.check(self._L.sadgpu_compute_nrgba(self._h, stream, l.ctypes.data, l.strides[0], r.ctypes.data, r.strides[0],
                                              w, h, block_size, max_disparity, out.ctypes.data, out.strides[0]))
        return out

    def submit(self, left, right, block_size, max_disparity, y0=0, y1=None, stream=0, out=None):
        """Enqueue one frame.  With `out` (an array from host_array) the result is written there by the D2H copy
        itself (sadgpu_submit_into) and wait(ticket) needs no destination."""
        l, r = _pair(left, right)
        h, w = l.shape
        y1 = h if y1 is None else y1
        t = ctypes.c_uint64()
        if out is not None:
            out = _out_2d(out, (h, w))
        if out is None:
            N.check(self._L.sadgpu_submit(self._h, stream, l.ctypes.data, l.strides[0], r.ctypes.data, r.strides[0],
                                           w, h, block_size, max_disparity, y0, y1, ctypes.byref(t)))
        else:
            N.check(self._L.sadgpu_submit_into(self._h, stream, l.ctypes.data, l.strides[0], r.ctypes.data, r.strides[0],
                                                w, h, block_size, max_disparity, y0, y1, out.ctypes.data, out.strides[0],
                                                ctypes.byref(t)))
        return t.value

    def wait_uploaded(self, ticket):
        """Blocks until the H2D copies of a submitted frame have run (pinned-pool inputs may be overwritten again)."""
        N.check(self._L.sadgpu_wait_uploaded(self._h, ticket))

    def wait(self, ticket, out=None):
        if out is not None:
            if not isinstance(out, np.ndarray) or out.dtype != np.uint8 or out.ndim != 2 or out.strides[1] != 1:
                raise ValueError("out: expected a uint8 plane with contiguous rows")
        if out is None:
            N.check(self._L.sadgpu_wait(self._h, ticket, None, 0))
        else:
            N.check(self._L.sadgpu_wait(self._h, ticket, out.ctypes.data, out.strides[0]))
        return out

    def reserve_batch(self, max_frames):
        N.check(self._L.sadgpu_reserve_batch(self._h, max_frames))

    def submit_batch(self, pairs, block_size, max_disparity, out, stream=0):
        """pairs: uint8 [n][2][h][w] (contiguous; pinned pool memory avoids the staging copy); out: pool array [n][h][w]."""
        p = np.asarray(pairs)
        if p.dtype != np.uint8 or p.ndim != 4 or p.shape[1] != 2 or not p.flags.c_contiguous:
            raise ValueError("pairs: expected a contiguous uint8 array [n][2][h][w]")
        n, _, h, w = p.shape
        if not isinstance(out, np.ndarray) or out.dtype != np.uint8 or out.shape != (n, h, w) or not out.flags.c_contiguous:
            raise ValueError("out: expected a contiguous uint8 array [n][h][w]")
        t = ctypes.c_uint64()
        N.check(self._L.sadgpu_submit_batch_into(self._h, stream, n, p.ctypes.data, w, h, block_size, max_disparity,
                                                  out.ctypes.data, ctypes.byref(t)))
        return t.value

    def compute_sharded(self, left, right, block_size, max_disparity, out=None):
        l, r = _pair(left, right)
        h, w = l.shape
        out = _out_2d(out, (h, w))
        N.check(self._L.sadgpu_compute_sharded(self._h, l.ctypes.data, l.strides[0], r.ctypes.data, r.strides[0],
                                                w, h, block_size, max_disparity, out.ctypes.data, out.strides[0]))
        return out

    # -- device-resident entry point -----------------------------------------------------
    def compute_device(self, dL, pitch_l, dR, pitch_r, w, h, block_size, max_disparity, dOut, pitch_out,
                       y0=0, y1=None, device=0, cuda_stream=0, tuning=None):
        y1 = h if y1 is None else y1
        tp = None
        if tuning:
            t = N.Tuning(); [setattr(t, k, v) for k, v in tuning.items()]
            tp = ctypes.byref(t)
        N.check(self._L.sadgpu_compute_device(self._h, device, dL, pitch_l, dR, pitch_r, w, h, block_size,
                                               max_disparity, y0, y1, dOut, pitch_out, cuda_stream, tp))

    def compute_device_batch(self, n_frames, dL, pitch_l, fs_l, dR, pitch_r, fs_r, w, h, block_size, max_disparity,
                             dOut, pitch_out, fs_out, y0=0, y1=None, device=0, cuda_stream=0, tuning=None):
        y1 = h if y1 is None else y1
        tp = None
        if tuning:
            t = N.Tuning(); [setattr(t, k, v) for k, v in tuning.items()]
            tp = ctypes.byref(t)
        N.check(self._L.sadgpu_compute_device_batch(self._h, device, n_frames, dL, pitch_l, fs_l, dR, pitch_r, fs_r,
                                                     w, h, block_size, max_disparity, y0, y1, dOut, pitch_out, fs_out,
                                                     cuda_stream, tp))

    def gray_device(self, dSrc, src_pitch, channels, mode, w, h, dGray, gray_pitch, device=0, cuda_stream=0):
        """Go-exact luma on the device (modes: 0 NRGBA generic path, 1 opaque RGB intended, 2 LoadPNG-as-written)."""
        N.check(self._L.sadgpu_gray_device(self._h, device, dSrc, src_pitch, channels, mode, w, h, dGray, gray_pitch, cuda_stream))

    # -- post-processing hooks (additive) -------------------------------
    def compute_checked(self, left, right, block_size, max_disparity, tolerance=1, invalid_value=0, median=False, stream=0, out=None):
        """Left-right consistency checked (and optionally 3x3 median filtered) disparity map; NOT part of the bit-exact path."""
        l, r = _pair(left, right)
        h, w = l.shape
        out = _out_2d(out, (h, w))
        N.check(self._L.sadgpu_compute_checked(self._h, stream, l.ctypes.data, l.strides[0], r.ctypes.data, r.strides[0], w, h,
                                                block_size, max_disparity, tolerance, invalid_value, int(bool(median)),
                                                out.ctypes.data, out.strides[0]))
        return out

    def median3_device(self, dSrc, src_pitch, w, h, dDst, dst_pitch, device=0, cuda_stream=0):
        N.check(self._L.sadgpu_median3_device(self._h, device, dSrc, src_pitch, w, h, dDst, dst_pitch, cuda_stream))

    def lrcheck_device(self, dLeft, left_pitch, dRight, right_pitch, w, h, max_disparity, tolerance, invalid_value, dDst, dst_pitch,
                       device=0, cuda_stream=0):
        N.check(self._L.sadgpu_lrcheck_device(self._h, device, dLeft, left_pitch, dRight, right_pitch, w, h, max_disparity, tolerance,
                                               invalid_value, dDst, dst_pitch, cuda_stream))

    # -- pinned pool -----------------------------------------------------------------------
    def host_array(self, shape):
        n = int(np.prod(shape))
        p = self._L.sadgpu_host_alloc(self._h, n)
        if not p:
            raise MemoryError("sadgpu_host_alloc failed")
        buf = (ctypes.c_uint8 * n).from_address(p)
        a = np.frombuffer(buf, np.uint8).reshape(shape)
        return a

    def host_pair(self, h, w):
        """left/right frames back to back in one pinned allocation: uploaded with a single DMA."""
        a = self.host_array((2, h, w))
        return a[0], a[1]

    def last_launch_count(self):
        return self._L.sadgpu_last_launch_count(self._h)


def plan_describe(w, h, block_size, max_disparity, y0=0, y1=None, tuning=None, frames=1):
    y1 = h if y1 is None else y1
    buf = ctypes.create_string_buffer(1024)
    t = N.Tuning()
    for k, v in (tuning or {}).items():
        setattr(t, k, v)
    t.reserved[0] = frames
    tp = ctypes.byref(t)
    N.check(N.lib().sadgpu_plan_describe(w, h, block_size, max_disparity, y0, y1, tp, buf, 1024))
    return json.loads(buf.value.decode())
