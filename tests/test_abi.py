"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/sadgpu.h
declares, plans launches, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def native():
    import importlib.util
    spec = importlib.util.spec_from_file_location("sadgpu_build", os.path.join(ROOT, "steroscopic-hardware_b200", "build.py"))
    b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
    b.build_all()
    from despair import _native
    return _native


def test_header_symbols_all_exported(native):
    hdr = open(os.path.join(ROOT, "include", "sadgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(sadgpu_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = ctypes.CDLL(native.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(native.EXPORTS), (declared ^ set(native.EXPORTS))


def test_version_and_strerror(native):
    L = native.lib()
    assert b"sm_100a" in L.sadgpu_version()
    assert b"invalid" in L.sadgpu_strerror(-1)
    assert b"no CPU fallback" in L.sadgpu_strerror(-5)


def test_plans_cover_the_parameter_surface(native):
    import despair
    for B in range(1, 32):
        for D in (1, 16, 32, 64, 68, 128, 255, 256):
            p = despair.plan_describe(1920, 1080, B, D)
            assert p["half"] == B // 2 and p["smem"] <= 232448 and p["NC"] * p["NGc"] >= p["NG"]
            assert p["grid"][0] * p["TW"] >= 1920
            if p["variant"] == "warp-specialised, shared-memory ring":
                assert (B >= 18 or (B >= 16 and D <= 32)) and (p["NGc"], p["TW"], p["RB"]) in ((9, 32, 10), (13, 32, 7))
                assert p["NGc"] == 9 or (B >= 18 and p["NG"] in (64, 65))             # 13-group chunks: D = 255, 256 among the D's of this loop
            elif p["variant"] == "warp-specialised":
                if B <= 9:
                    assert (p["NGc"], p["TW"]) in ((33, 32), (17, 64), (9, 96), (5, 192))
                    assert p["NC"] == 1 or p["NGc"] == 33            # only the 33-group layout ever chunks the range
                else:
                    assert B <= 17 and p["RB"] == 2 * (B // 2) + 2
                    assert (p["NGc"], p["TW"]) == ((9, 64) if (D <= 32 and B <= 15) else (17, 32))
            else:
                assert p["variant"] == "fast" and 10 <= B <= 15 and p["NG"] == 18
    assert despair.plan_describe(1920, 1080, 9, 128)["variant"] == "warp-specialised"
    assert despair.plan_describe(640, 480, 9, 64)["NGc"] == 17                    # cfg1: two strips x 16 groups + tail
    assert despair.plan_describe(1920, 1080, 9, 16)["NGc"] == 5
    assert despair.plan_describe(1920, 1080, 15, 256)["variant"] == "warp-specialised"     # cfg2: four chunks of 17 groups
    assert despair.plan_describe(1920, 1080, 15, 68)["variant"] == "fast"
    assert despair.plan_describe(1920, 1080, 31, 256)["variant"] == "warp-specialised, shared-memory ring"   # cfg4: five chunks of 13 groups
    assert (despair.plan_describe(1920, 1080, 31, 256)["NGc"], despair.plan_describe(1920, 1080, 31, 48)["NGc"]) == (13, 13)
    assert (despair.plan_describe(1920, 1080, 31, 64)["NGc"], despair.plan_describe(1920, 1080, 31, 128)["NGc"]) == (9, 9)
    assert despair.plan_describe(1920, 1080, 31, 16)["variant"] == "warp-specialised, shared-memory ring"
    assert despair.plan_describe(1920, 1080, 16, 16)["variant"] == "warp-specialised, shared-memory ring"
    assert despair.plan_describe(1920, 1080, 31, 256, tuning=dict(kernel_variant=6))["variant"] == "ring"
    assert despair.plan_describe(1920, 1080, 31, 16, tuning=dict(kernel_variant=4))["variant"] == "wide"
    assert despair.plan_describe(1920, 1080, 16, 64)["variant"] == "warp-specialised"     # the reference's start-up parameters (params.go:13-18)
    with pytest.raises(despair.SadGpuError):
        despair.plan_describe(1920, 1080, 31, 256, tuning=dict(kernel_variant=2))      # phase-alternating kernel needs block_size <= 15
    with pytest.raises(despair.SadGpuError):
        despair.plan_describe(1920, 1080, 19, 64, tuning=dict(kernel_variant=3))       # warp-specialised kernel needs block_size <= 17
    with pytest.raises(despair.SadGpuError):
        despair.plan_describe(1920, 1080, 9, 64, tuning=dict(kernel_variant=7))        # its shared-memory-ring form needs block_size >= 10
    for gone in (1, 5):                                                               # removed kernels are rejected, not substituted
        with pytest.raises(despair.SadGpuError):
            despair.plan_describe(1920, 1080, 15, 64, tuning=dict(kernel_variant=gone))


def test_plan_rejects_bad_parameters(native):
    import despair
    for (w, h, B, D, y0, y1) in [(0, 10, 9, 64, 0, 10), (10, 10, 0, 64, 0, 10), (10, 10, 32, 64, 0, 10),
                                 (10, 10, 9, 0, 0, 10), (10, 10, 9, 257, 0, 10)]:
        with pytest.raises(despair.SadGpuError) as e:
            despair.plan_describe(w, h, B, D, y0, y1)
        assert e.value.code == -1
    with pytest.raises(despair.SadGpuError) as e:
        despair.plan_describe(10, 10, 9, 64, 5, 11)
    assert e.value.code == -2


def test_no_cpu_fallback_without_gpu(native):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import despair
    assert native.lib().sadgpu_device_count() == 0
    with pytest.raises(despair.SadGpuError):
        despair.Context([0], 64, 64, 1)


def test_missing_library_fails_loudly(native, monkeypatch):
    monkeypatch.setattr(native, "_lib", None)
    monkeypatch.setattr(native, "LIB_PATH", "/nonexistent/libsadgpu.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        native.lib()
