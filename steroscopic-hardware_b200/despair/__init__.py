"""Host-side mirror of the reference package pkg/despair over libsadgpu.so (B200, sm_100a).

Nothing here computes disparities on the CPU; every call ends in the C ABI of include/sadgpu.h.
"""
from ._native import SadGpuError, lib, LIB_PATH  # noqa: F401
from .context import Context, plan_describe  # noqa: F401
