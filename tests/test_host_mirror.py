"""C++ mirror of pkg/despair (steroscopic-hardware_b200/host): planner logic on CPU, the pipeline on GPU."""
import ctypes
import os

import numpy as np
import pytest

from conftest import ROOT

HOST_LIB = os.path.join(ROOT, "steroscopic-hardware_b200", "libdespair_host.so")


@pytest.fixture(scope="module")
def host():
    import importlib.util
    spec = importlib.util.spec_from_file_location("sadgpu_build", os.path.join(ROOT, "steroscopic-hardware_b200", "build.py"))
    b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
    b.build_all()
    L = ctypes.CDLL(HOST_LIB)
    u8p = ctypes.c_void_p
    L.despair_host_run_sad.argtypes = [u8p, u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, u8p]
    L.despair_host_pipeline.argtypes = [u8p, u8p] + [ctypes.c_int] * 7 + [u8p, ctypes.POINTER(ctypes.c_int)]
    L.despair_host_stream.argtypes = [u8p, u8p] + [ctypes.c_int] * 6 + [u8p]
    L.despair_host_run_sad_chunks.argtypes = [ctypes.c_int] * 3 + [ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    return L


def test_run_sad_chunk_planner_matches_reference_restatement(host, oracle):
    """RunSadChunks == the Python restatement of pkg/despair/sad.go:128-153 for several image sizes / core counts."""
    for (w, h, ncpu) in [(640, 480, 8), (1920, 1080, 16), (3840, 2160, 64), (100, 37, 1), (333, 77, 2)]:
        buf = (ctypes.c_int * (4 * 8192))()
        n = host.despair_host_run_sad_chunks(w, h, ncpu, buf, 8192)
        exp = oracle.run_sad_chunks(w, h, ncpu)
        assert n == len(exp)
        got = [tuple(buf[4 * i:4 * i + 4]) for i in range(n)]
        assert got == exp


def test_planner_error_cases_match_go_panics(host):
    buf = (ctypes.c_int * 16)()
    assert host.despair_host_run_sad_chunks(4, 4, 8, buf, 4) == -1       # W*H < numChunks -> divide by zero in Go


@pytest.mark.gpu
def test_run_sad_on_gpu(host, oracle):
    rng = np.random.default_rng(3)
    base = rng.integers(0, 256, (120, 260), dtype=np.uint8)
    L = np.ascontiguousarray(base[:, 60:]); R = np.ascontiguousarray(np.roll(base, -17, 1)[:, 60:])
    out = np.zeros_like(L)
    assert host.despair_host_run_sad(L.ctypes.data, R.ctypes.data, 200, 120, 16, 64, out.ctypes.data) == 0
    assert np.array_equal(out, oracle.frame_box(L, R, 16, 64))


@pytest.mark.gpu
def test_stream_sad_video_path_on_gpu(host, oracle):
    """StreamSad: pinned frame pairs in, pinned maps out, several pairs per GPU call (examples/run.stream.go:33-67)."""
    rng = np.random.default_rng(6)
    n, h, w = 11, 96, 160
    base = rng.integers(0, 256, (n, h, w + 64), dtype=np.uint8)
    L = np.ascontiguousarray(base[:, :, 64:]); R = np.ascontiguousarray(np.roll(base, -9, 2)[:, :, 64:])
    out = np.zeros_like(L)
    assert host.despair_host_stream(L.ctypes.data, R.ctypes.data, n, w, h, 11, 96, 4, out.ctypes.data) == 0
    for i in range(n):
        assert np.array_equal(out[i], oracle.frame_box(L[i], R[i], 11, 96)), i


@pytest.mark.gpu
def test_output_camera_style_pipeline_on_gpu(host, oracle):
    """SetupConcurrentSAD(32) + H/128-row bands + AssembleDisparityMap, as pkg/camera/output.go:172-190 drives it."""
    rng = np.random.default_rng(4)
    h, w = 480, 640
    base = rng.integers(0, 256, (h, w + 64), dtype=np.uint8)
    L = np.ascontiguousarray(base[:, 64:]); R = np.ascontiguousarray(np.roll(base, -21, 1)[:, 64:])
    exp = oracle.frame_box(L, R, 9, 64)
    out = np.zeros_like(L); seen = (ctypes.c_int * 2)()
    assert host.despair_host_pipeline(L.ctypes.data, R.ctypes.data, w, h, 9, 64, 32, max(1, h // 128), 0, out.ctypes.data, seen) == 0
    assert np.array_equal(out, exp) and tuple(seen) == (9, 64)
    # faithful mode reproduces sad.go:179-184: exactly one band (the last to arrive) stays zero
    out2 = np.zeros_like(L)
    assert host.despair_host_pipeline(L.ctypes.data, R.ctypes.data, w, h, 9, 64, 32, 3, 1, out2.ctypes.data, None) == 0
    diff_rows = np.nonzero((out2 != exp).any(axis=1))[0]
    assert len(diff_rows) <= 3 and (len(diff_rows) == 0 or (diff_rows.max() - diff_rows.min() < 3 and not out2[diff_rows].any()))
    host.despair_host_shutdown()
