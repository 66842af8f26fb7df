"""ctypes binding of libsadgpu.so (include/sadgpu.h).  Loading fails loudly when the CUDA
library has not been built: there is no CPU fallback anywhere in this package."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libsadgpu.so")

u8p = ctypes.POINTER(ctypes.c_uint8)
c_int = ctypes.c_int
c_size_t = ctypes.c_size_t
c_void_p = ctypes.c_void_p


class Tuning(ctypes.Structure):
    _fields_ = [("rows_per_batch", c_int), ("band_rows", c_int), ("groups_per_chunk", c_int),
                ("kernel_variant", c_int), ("reserved", c_int * 4)]


EXPORTS = {
    "sadgpu_device_count": (c_int, []),
    "sadgpu_create": (c_int, [ctypes.POINTER(c_int), c_int, c_int, c_int, c_int, ctypes.POINTER(c_void_p)]),
    "sadgpu_destroy": (None, [c_void_p]),
    "sadgpu_compute": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                               c_int, c_int, c_void_p, c_int]),
    "sadgpu_compute_region": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                      c_int, c_int, c_int, c_int, c_void_p, c_int]),
    "sadgpu_region_stats": (c_int, [c_void_p, ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_longlong),
                                    ctypes.POINTER(ctypes.c_longlong)]),
    "sadgpu_wait_uploaded": (c_int, [c_void_p, ctypes.c_uint64]),
    "sadgpu_compute_nrgba": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                     c_void_p, c_int]),
    "sadgpu_submit": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                              c_int, c_int, ctypes.POINTER(ctypes.c_uint64)]),
    "sadgpu_wait": (c_int, [c_void_p, ctypes.c_uint64, c_void_p, c_int]),
    "sadgpu_submit_into": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                   c_int, c_int, c_void_p, c_int, ctypes.POINTER(ctypes.c_uint64)]),
    "sadgpu_reserve_batch": (c_int, [c_void_p, c_int]),
    "sadgpu_submit_batch_into": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                         ctypes.POINTER(ctypes.c_uint64)]),
    "sadgpu_compute_sharded": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                       c_void_p, c_int]),
    "sadgpu_compute_device": (c_int, [c_void_p, c_int, c_void_p, c_size_t, c_void_p, c_size_t, c_int, c_int,
                                      c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p,
                                      ctypes.POINTER(Tuning)]),
    "sadgpu_compute_device_batch": (c_int, [c_void_p, c_int, c_int, c_void_p, c_size_t, c_size_t, c_void_p, c_size_t,
                                            c_size_t, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t,
                                            c_size_t, c_void_p, ctypes.POINTER(Tuning)]),
    "sadgpu_gray_device": (c_int, [c_void_p, c_int, c_void_p, c_size_t, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "sadgpu_median3_device": (c_int, [c_void_p, c_int, c_void_p, c_size_t, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "sadgpu_lrcheck_device": (c_int, [c_void_p, c_int, c_void_p, c_size_t, c_void_p, c_size_t, c_int, c_int, c_int, c_int, c_int,
                                      c_void_p, c_size_t, c_void_p]),
    "sadgpu_compute_checked": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                       c_int, c_void_p, c_int]),
    "sadgpu_host_alloc": (c_void_p, [c_void_p, c_size_t]),
    "sadgpu_host_free": (None, [c_void_p, c_void_p]),
    "sadgpu_debug_read": (c_int, [c_void_p, c_int, ctypes.POINTER(ctypes.c_uint32), c_int]),
    "sadgpu_last_launch_count": (c_int, [c_void_p]),
    "sadgpu_plan_describe": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(Tuning),
                                     ctypes.c_char_p, c_size_t]),
    "sadgpu_strerror": (ctypes.c_char_p, [c_int]),
    "sadgpu_version": (ctypes.c_char_p, []),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python steroscopic-hardware_b200/build.py` "
                "(nvcc, sm_100a).  The SAD path has no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(L, name)          # AttributeError == symbol missing from the library
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class SadGpuError(RuntimeError):
    def __init__(self, code):
        self.code = code
        super().__init__(f"sadgpu error {code}: {lib().sadgpu_strerror(code).decode()}")


def check(code):
    if code != 0:
        raise SadGpuError(code)
