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
    L.despair_host_output_camera_loop.argtypes = [u8p, u8p] + [ctypes.c_int] * 9 + [u8p, ctypes.POINTER(ctypes.c_double)]
    L.despair_host_process_depth_map.argtypes = [u8p, u8p] + [ctypes.c_int] * 4 + [u8p]
    L.despair_host_serial_stream.argtypes = [u8p, u8p, ctypes.c_size_t] + [ctypes.c_int] * 5 + [ctypes.c_uint, u8p]
    L.despair_host_configure.argtypes = [ctypes.c_int] * 3
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


@pytest.mark.gpu
def test_output_camera_frame_loop_on_gpu(host, oracle):
    """The OutputCamera loop (pkg/camera/output.go:129-210 minus PNG files): one 32-worker pipeline, several frames, H/128-row
    bands.  Fresh image objects per frame (the reference's image.NewGray) and the same two objects refilled in place."""
    rng = np.random.default_rng(41)
    h, w, n = 480, 640, 3
    base = rng.integers(0, 256, (n, h, w + 64), dtype=np.uint8)
    L = np.ascontiguousarray(base[:, :, 64:]); R = np.ascontiguousarray(np.roll(base, -19, 2)[:, :, 64:])
    for reuse in (0, 1):
        out = np.zeros((h, w), np.uint8); us = ctypes.c_double()
        iters, warm = 5, 2
        assert host.despair_host_output_camera_loop(L.ctypes.data, R.ctypes.data, n, w, h, 16, 64, 32, warm, iters, reuse,
                                                    out.ctypes.data, ctypes.byref(us)) == 0
        k = (warm + iters - 1) % n
        assert np.array_equal(out, oracle.frame_box(L[k], R[k], 16, 64)), reuse
        assert us.value > 0
    host.despair_host_shutdown()


@pytest.mark.gpu
def test_backend_errors_are_reported_not_blanked(host):
    """Round-1 advisor finding: a failing backend call must not come back as rc 0 with an all-zero map."""
    L = np.zeros((64, 64), np.uint8); out = np.zeros_like(L)
    assert host.despair_host_run_sad(L.ctypes.data, L.ctypes.data, 64, 64, 33, 64, out.ctypes.data) == -1     # block size > 31
    assert host.despair_host_pipeline(L.ctypes.data, L.ctypes.data, 64, 64, 40, 64, 4, 8, 0, out.ctypes.data, None) == -2
    host.despair_host_configure(32, 32, 2)                                       # image larger than the backend
    assert host.despair_host_pipeline(L.ctypes.data, L.ctypes.data, 64, 64, 9, 64, 4, 8, 0, out.ctypes.data, None) == -2
    host.despair_host_configure(4096, 2304, 4)
    assert host.despair_host_pipeline(L.ctypes.data, L.ctypes.data, 64, 64, 9, 64, 4, 8, 0, out.ctypes.data, None) == 0
    host.despair_host_shutdown()


@pytest.mark.gpu
def test_process_depth_map_from_decoded_colour_pair(host, oracle, manifest):
    """N2: ProcessDepthMap takes the decoded NRGBA pair (what png.Decode returns for the reference's testdata), the luma is
    taken on the device with Go's arithmetic, no PNG round trip (pkg/camera/output.go:129-210)."""
    from PIL import Image
    from conftest import GOLDEN
    from oracle.go_image import load_png
    m = manifest["rgba_crop"]
    pl = os.path.join(GOLDEN, f"L_{m['tag']}_rgba_crop.png"); pr = os.path.join(GOLDEN, f"R_{m['tag']}_rgba_crop.png")
    L4 = np.ascontiguousarray(np.array(Image.open(pl))); R4 = np.ascontiguousarray(np.array(Image.open(pr)))
    gl = load_png(pl, "intended"); gr = load_png(pr, "intended")
    out = np.zeros((m["h"], m["w"]), np.uint8)
    for (B, D) in ((16, 64), (9, 64)):
        assert host.despair_host_process_depth_map(L4.ctypes.data, R4.ctypes.data, m["w"], m["h"], B, D, out.ctypes.data) == 0
        assert np.array_equal(out, oracle.frame_box(gl, gr, B, D)), (B, D)
    host.despair_host_shutdown()


@pytest.mark.gpu
def test_serial_ingest_into_pinned_frames(host, oracle):
    """N3: two raw-gray byte streams delivered in ragged reads of 1..1024 bytes (pkg/camera/serial.go:274-295) land directly
    in pinned frame pairs; frame k+1 is read while frame k is on the GPU; a stream that ends mid-frame ends cleanly."""
    rng = np.random.default_rng(42)
    n, h, w = 5, 120, 256
    base = rng.integers(0, 256, (n, h, w + 64), dtype=np.uint8)
    L = np.ascontiguousarray(base[:, :, 64:]); R = np.ascontiguousarray(np.roll(base, -11, 2)[:, :, 64:])
    out = np.zeros((n, h, w), np.uint8)
    got = host.despair_host_serial_stream(L.ctypes.data, R.ctypes.data, L.size, w, h, 9, 64, 100, 7, out.ctypes.data)
    assert got == n
    for i in range(n):
        assert np.array_equal(out[i], oracle.frame_box(L[i], R[i], 9, 64)), i
    out[:] = 0                                                   # the right camera stops in the middle of its fourth frame
    got = host.despair_host_serial_stream(L.ctypes.data, R.ctypes.data, 3 * h * w + 1000, w, h, 16, 64, 100, 9, out.ctypes.data)
    assert got == 3 and not out[3:].any()
    for i in range(3):
        assert np.array_equal(out[i], oracle.frame_box(L[i], R[i], 16, 64)), i
    host.despair_host_shutdown()
