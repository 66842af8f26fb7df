"""GPU parity tests proper: libsadgpu.so (through its C ABI) against the oracle and the golden
fixtures.  Bit-exact: this is integer/byte work."""
import hashlib
import os

import numpy as np
import pytest

from conftest import load_gray, GOLDEN

pytestmark = pytest.mark.gpu
sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def ctx():
    import despair
    c = despair.Context([0], 3840, 2160, 4)
    yield c
    c.close()


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    assert torch.cuda.is_available()
    return torch


def dev_run(torch, ctx, L, R, B, D, tuning=None, y0=0, y1=None, pitch_pad=0):
    h, w = L.shape
    pl = w + pitch_pad
    dL = torch.zeros((h, pl), dtype=torch.uint8, device="cuda"); dL[:, :w] = torch.from_numpy(L).cuda()
    dR = torch.zeros((h, pl), dtype=torch.uint8, device="cuda"); dR[:, :w] = torch.from_numpy(R).cuda()
    dO = torch.full((h, pl), 77, dtype=torch.uint8, device="cuda")
    ctx.compute_device(dL.data_ptr(), pl, dR.data_ptr(), pl, w, h, B, D, dO.data_ptr(), pl, y0=y0, y1=y1,
                       cuda_stream=torch.cuda.current_stream().cuda_stream, tuning=tuning)
    torch.cuda.synchronize()
    return dO.cpu().numpy()[:, :w]


def synth_pair(rng, H, W, kind):
    if kind == 0:
        return rng.integers(0, 256, (H, W), dtype=np.uint8), rng.integers(0, 256, (H, W), dtype=np.uint8)
    if kind == 1:   # shifted texture: a true disparity exists
        base = rng.integers(0, 256, (H, W + 64), dtype=np.uint8); s = int(rng.integers(0, 40))
        return base[:, 64:64 + W].copy(), np.roll(base, -s, 1)[:, 64:64 + W].copy()
    if kind == 2:   # three grey levels: massive ties
        return rng.integers(0, 3, (H, W), dtype=np.uint8), rng.integers(0, 3, (H, W), dtype=np.uint8)
    if kind == 3:   # maximum SAD everywhere (255 vs 0): exercises the 16-bit headroom and the poison ordering
        return np.full((H, W), 255, np.uint8), np.zeros((H, W), np.uint8)
    return np.full((H, W), 31, np.uint8), np.full((H, W), 31, np.uint8)   # flat: all candidates tie


@pytest.mark.parametrize("B", list(range(1, 32)))
def test_every_block_size_random(torch_mod, ctx, oracle, B):
    rng = np.random.default_rng(100 + B)
    for i, D in enumerate((16, 64, 100, 256)):
        W = int(rng.integers(40, 220)); H = int(rng.integers(20, 80))
        L, R = synth_pair(rng, H, W, i % 5)
        assert np.array_equal(dev_run(torch_mod, ctx, L, R, B, D), oracle.frame_box(L, R, B, D)), (W, H, B, D)


@pytest.mark.parametrize("D", [1, 2, 3, 4, 5, 15, 16, 17, 32, 48, 64, 127, 128, 129, 255, 256])
def test_every_disparity_grid_point(torch_mod, ctx, oracle, D):
    rng = np.random.default_rng(200 + D)
    for i, B in enumerate((3, 9, 16, 31)):
        W = int(rng.integers(30, 300)); H = int(rng.integers(10, 60))
        L, R = synth_pair(rng, H, W, (i + D) % 5)
        assert np.array_equal(dev_run(torch_mod, ctx, L, R, B, D), oracle.frame_box(L, R, B, D)), (W, H, B, D)


def test_degenerate_shapes(torch_mod, ctx, oracle):
    rng = np.random.default_rng(5)
    for (W, H, B, D) in [(1, 1, 1, 1), (1, 1, 31, 256), (2, 50, 9, 64), (50, 2, 9, 64), (5, 5, 31, 256),
                         (7, 40, 15, 16), (300, 3, 3, 256), (64, 64, 9, 200), (57, 33, 8, 20), (113, 17, 30, 77)]:
        for kind in (0, 3, 4):
            L, R = synth_pair(rng, H, W, kind)
            assert np.array_equal(dev_run(torch_mod, ctx, L, R, B, D), oracle.frame_box(L, R, B, D)), (W, H, B, D, kind)


def test_tilings_do_not_change_pixels(torch_mod, ctx, oracle):
    """Forced small batches / bands / disparity chunks: every CTA boundary is crossed."""
    rng = np.random.default_rng(6)
    for i in range(40):
        W = int(rng.integers(20, 260)); H = int(rng.integers(10, 120))
        B = int(rng.integers(1, 32)); D = int(rng.choice([5, 16, 64, 128, 256]))
        L, R = synth_pair(rng, H, W, i % 5)
        if B > 17 or (B > 15 and i % 2 == 0):
            tun = dict(band_rows=int(rng.integers(1, 50)), kernel_variant=4)
        elif B <= 17 and i % 2:         # warp-specialised kernel: forced chunk sizes 33 / 17 / 9 / 5 groups = every lane layout, chunked ranges
            tun = dict(band_rows=int(rng.integers(1, 50)), groups_per_chunk=int(rng.choice([33, 17, 9, 5])), kernel_variant=3)
        else:
            tun = dict(band_rows=int(rng.integers(1, 50)), groups_per_chunk=int(rng.integers(1, 21)), kernel_variant=2)
        assert np.array_equal(dev_run(torch_mod, ctx, L, R, B, D, tun), oracle.frame_box(L, R, B, D)), (W, H, B, D, tun)
    # the barrier-pipelined kernels: bands shorter than a burst / the window, one-row bands, bands that end mid-burst
    for i in range(40):
        W = int(rng.integers(20, 260)); H = int(rng.integers(10, 120))
        B = int(rng.integers(10, 32)); D = int(rng.choice([5, 16, 64, 128, 256]))
        L, R = synth_pair(rng, H, W, i % 5)
        tun = dict(band_rows=int(rng.integers(1, 50)), kernel_variant=6)
        assert np.array_equal(dev_run(torch_mod, ctx, L, R, B, D, tun), oracle.frame_box(L, R, B, D)), (W, H, B, D, tun)


@pytest.mark.parametrize("variant", [2, 3, 4, 6])
def test_every_kernel_variant_agrees_with_oracle(torch_mod, ctx, oracle, variant):
    """variant 2 = phase-alternating register-ring kernel (B <= 15), 3 = warp-specialised double-buffered kernel (B <= 17, every D:
    1 / 2 / 3 / 6 strips per CTA at B <= 9, 17-group chunks with 16- and 32-bit sums at B 10..17), 4 = large-window kernel
    (B 16..31), 6 = H-ring mbarrier-pipelined kernel (B 10..31)."""
    rng = np.random.default_rng(60 + variant)
    for i in range(48 if variant == 3 else 24):
        W = int(rng.integers(20, 400)); H = int(rng.integers(10, 100))
        if variant == 4:
            B = int(rng.integers(16, 32)); D = int(rng.choice([7, 16, 33, 64, 128, 200, 256]))
        elif variant == 6:
            B = int(rng.integers(10, 32)); D = int(rng.choice([7, 16, 33, 64, 128, 200, 256]))
        elif variant == 3:
            B = int(rng.integers(1, 18)); D = int(rng.choice([1, 7, 16, 17, 20, 32, 33, 36, 48, 64, 65, 68, 100, 128, 129, 200, 256]))
        else:
            B = int(rng.integers(1, 16)); D = int(rng.choice([7, 16, 33, 64, 128, 200, 256]))
        L, R = synth_pair(rng, H, W, i % 5)
        got = dev_run(torch_mod, ctx, L, R, B, D, dict(kernel_variant=variant))
        assert np.array_equal(got, oracle.frame_box(L, R, B, D)), (W, H, B, D, variant)


STRICT_WSR = dict(kernel_variant=7, reserved=(0, 0, 2, 0))       # reserved[2] = 2: fail instead of substituting another kernel


@pytest.mark.parametrize("B", list(range(10, 32)))
def test_shared_ring_warp_specialised_kernel(torch_mod, ctx, oracle, B):
    """variant 7 (the planner's choice for block_size >= 18): widths that are multiples of 16 so that TMA addresses the images, and
    the strict flag, so that a silent substitution of another kernel would fail the test.  One chunk (D <= 32), chunked ranges,
    forced bands shorter than a batch / the window, every image kind."""
    rng = np.random.default_rng(700 + B)
    for i, D in enumerate((16, 32, 36, 44, 48, 56, 60, 64, 100, 255, 256)):          # last chunk of 5, 9, 1, 3, 4, 6, 7, 8, 8, 1, 2 groups
        W = 16 * int(rng.integers(2, 24)); H = int(rng.integers(8, 90))
        L, R = synth_pair(rng, H, W, (i + B) % 5)
        tun = dict(STRICT_WSR)
        if i % 2:
            tun["band_rows"] = int(rng.integers(1, 40))
        if B >= 18:
            tun["groups_per_chunk"] = 9 if (i + B) % 2 else 13                 # both chunk sizes (9 groups x 10 rows, 13 groups x 7 rows)
        assert np.array_equal(dev_run(torch_mod, ctx, L, R, B, D, tun), oracle.frame_box(L, R, B, D)), (W, H, B, D, tun)


def test_shared_ring_kernel_row_ranges_frames_and_fallback(torch_mod, ctx, oracle):
    import despair
    torch = torch_mod
    rng = np.random.default_rng(77)
    # row ranges [y0, y1) of a frame (RunSad tiles / row-band sharding)
    L, R = synth_pair(rng, 97, 208, 1)
    exp = oracle.frame_box(L, R, 25, 64)
    for (y0, y1) in ((0, 97), (0, 1), (13, 14), (5, 60), (60, 97), (96, 97)):
        got = dev_run(torch_mod, ctx, L, R, 25, 64, STRICT_WSR, y0=y0, y1=y1)
        assert np.array_equal(got[y0:y1], exp[y0:y1]) and (got[:y0] == 77).all() and (got[y1:] == 77).all(), (y0, y1)
    # several frames per launch (TMA walks the frame axis), one and several chunks
    for (F, H, W, B, D) in [(3, 41, 96, 31, 32), (4, 33, 160, 19, 200), (2, 70, 64, 16, 20)]:
        pairs = [synth_pair(rng, H, W, k % 5) for k in range(F)]
        dL = torch.from_numpy(np.stack([p[0] for p in pairs])).cuda(); dR = torch.from_numpy(np.stack([p[1] for p in pairs])).cuda()
        dO = torch.zeros_like(dL)
        ctx.compute_device_batch(F, dL.data_ptr(), W, W * H, dR.data_ptr(), W, W * H, W, H, B, D, dO.data_ptr(), W, W * H,
                                 cuda_stream=torch.cuda.current_stream().cuda_stream, tuning=STRICT_WSR)
        torch.cuda.synchronize()
        for f in range(F):
            assert np.array_equal(dO[f].cpu().numpy(), oracle.frame_box(pairs[f][0], pairs[f][1], B, D)), (F, H, W, B, D, f)
    # images TMA cannot address: the strict flag reports it, the default substitutes the ring / wide kernel (same pixels)
    L, R = synth_pair(rng, 40, 131, 0)
    with pytest.raises(despair.SadGpuError):
        dev_run(torch_mod, ctx, L, R, 31, 64, STRICT_WSR)
    for (B, D) in ((31, 64), (19, 16), (16, 32), (13, 16), (11, 128), (15, 68)):      # ring, wide and (block <= 15) phase-alternating fallbacks
        assert np.array_equal(dev_run(torch_mod, ctx, L, R, B, D, dict(kernel_variant=7)), oracle.frame_box(L, R, B, D)), (B, D)


def test_batched_frames_one_launch(torch_mod, ctx, oracle):
    torch = torch_mod
    rng = np.random.default_rng(61)
    for (F, H, W, B, D) in [(5, 60, 200, 9, 128), (3, 40, 130, 15, 256), (4, 33, 70, 31, 64)]:
        pairs = [synth_pair(rng, H, W, k % 5) for k in range(F)]
        dL = torch.from_numpy(np.stack([p[0] for p in pairs])).cuda(); dR = torch.from_numpy(np.stack([p[1] for p in pairs])).cuda()
        dO = torch.zeros_like(dL)
        ctx.compute_device_batch(F, dL.data_ptr(), W, W * H, dR.data_ptr(), W, W * H, W, H, B, D, dO.data_ptr(), W, W * H,
                                 cuda_stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        got = dO.cpu().numpy()
        for k in range(F):
            assert np.array_equal(got[k], oracle.frame_box(pairs[k][0], pairs[k][1], B, D)), (k, H, W, B, D)


def test_row_ranges_pitch_and_untouched_rows(torch_mod, ctx, oracle):
    rng = np.random.default_rng(7)
    L, R = synth_pair(rng, 90, 200, 1)
    exp = oracle.frame_box(L, R, 11, 64)
    got = dev_run(torch_mod, ctx, L, R, 11, 64, y0=17, y1=61, pitch_pad=24)
    assert np.array_equal(got[17:61], exp[17:61])
    assert (got[:17] == 77).all() and (got[61:] == 77).all()      # rows outside [y0,y1) are not written


def test_golden_cfg1_pairs_host_api(ctx, manifest):
    for tag, rec in manifest["pairs"].items():
        L = load_gray(f"L_{tag}_gray.png"); R = load_gray(f"R_{tag}_gray.png")
        got = ctx.compute(L, R, 9, 64)
        assert sha(got) == rec["b9_d64_sha256"] == manifest["survey_pins"][f"{tag}_b9_d64"]
        assert np.array_equal(got, load_gray(f"disp_{tag}_b9_d64.png"))
    L = load_gray("L_00001_gray.png"); R = load_gray("R_00001_gray.png")
    assert sha(ctx.compute(L, R, 16, 64)) == manifest["survey_pins"]["00001_b16_d64"]    # init default B=16 (params.go:13-18)


def test_golden_cfg2_native_resolution(ctx, manifest):
    L = load_gray("im0_intended_gray.png"); R = load_gray("im1_intended_gray.png")
    got = ctx.compute(L, R, 15, 256)
    assert sha(got) == manifest["survey_pins"]["im0_im1_intended_b15_d256"]
    z = np.zeros((1080, 1920), np.uint8)         # what LoadPNG as written feeds the path (gray.go:35-37)
    assert sha(ctx.compute(z, z, 15, 256)) == manifest["survey_pins"]["zeros_1920x1080"]


def test_fpga_known_answers_interior(ctx):
    v = np.load(os.path.join(GOLDEN, "fpga_vectors.npz"))
    for k in range(4):
        got = ctx.compute(v["L"][k], v["R"][k], 15, 64)
        exp = (v["exp_disp_p"][k].astype(int) * 255 // 64).astype(np.uint8)
        assert np.array_equal(got[:, 71:121], exp[:, 71:121])
    for k in range(v["hw_sad_L"].shape[0]):       # reference C generator run on random patches
        got = ctx.compute(v["hw_sad_L"][k], v["hw_sad_R"][k], 15, 64)
        exp = (v["hw_sad_out"][k].astype(int) * 255 // 64).astype(np.uint8)
        assert np.array_equal(got[7:121, 71:121], exp[7:121, 71:121])


def _stream_frame(seed, H=1080, W=1920, ramp=56):
    """cfg3/cfg4 generator of SURVEY.md §8(d)."""
    rng = np.random.default_rng(seed)
    T = rng.integers(0, 256, (H, W + 128 + 8 + ramp + 2), dtype=np.uint8).astype(np.uint16)
    T = ((T[:, :-2] + T[:, 1:-1] + T[:, 2:]) // 3).astype(np.uint8)     # light blur for texture
    delta = 8 + (ramp * np.arange(H)) // H
    L = T[:, 128:128 + W]
    idx = (np.arange(W)[None, :] + 128 + delta[:, None])
    R = np.take_along_axis(T, idx, axis=1)
    R = np.clip(R.astype(np.int16) + rng.integers(-2, 3, R.shape), 0, 255).astype(np.uint8)
    return np.ascontiguousarray(L), R


def test_cfg3_full_frame_1080p(ctx, oracle):
    L, R = _stream_frame(1234)
    got = ctx.compute(L, R, 9, 128)
    assert np.array_equal(got, oracle.frame_box(L, R, 9, 128))
    # the planted disparity ramp is recovered in the interior
    d = (got.astype(int) * 128 + 254) // 255
    assert abs(int(np.median(d[540, 300:1800])) - (8 + 56 * 540 // 1080)) <= 1


def test_cfg4_4k_b31_d256_rows_and_properties(ctx, oracle):
    L, R = _stream_frame(4321, 2160, 3840, 240)
    got = ctx.compute(L, R, 31, 256)
    for (a, b) in [(0, 24), (1000, 1016), (2140, 2160)]:
        assert np.array_equal(got[a:b], oracle.frame_box(L, R, 31, 256, a, b)), (a, b)
    assert not got[:, :15].any()                                           # X < h -> 0 (sad.go:212-218)
    x = np.arange(3840)
    assert (got.astype(int) <= (np.clip(x - 15, 0, 256) * 255) // 256).all()   # d <= X - h
    # row-range calls assemble the same map (chunk geometry never changes pixels)
    out = np.zeros_like(got)
    for i in range(8):
        ctx.compute(L, R, 31, 256, y0=270 * i, y1=270 * (i + 1), out=out, stream=i % 4)
    assert np.array_equal(out, got)
    # vertical-shift equivariance away from the borders
    got2 = ctx.compute(L[40:], R[40:], 31, 256)
    assert np.array_equal(got2[15:-15], got[55:-15])


@pytest.mark.parametrize("B,D", [(9, 64), (9, 128), (15, 128), (16, 64)])
def test_submit_wait_pipelining_and_sharded(ctx, oracle, B, D):
    """Four streams in flight (kernels of different frames overlap on the device) for every planner default:
    phase-alternating, warp-specialised + TMA, H-ring with 33- and 17-group chunks."""
    rng = np.random.default_rng(9)
    frames = [synth_pair(rng, 120, 320, 1) for _ in range(8)]
    exp = [oracle.frame_box(L, R, B, D) for L, R in frames]
    outs = [np.zeros((120, 320), np.uint8) for _ in frames]
    tickets = {}
    for k, (L, R) in enumerate(frames):
        s = k % 4
        if s in tickets:
            kk, t = tickets.pop(s); ctx.wait(t, outs[kk])
        tickets[s] = (k, ctx.submit(L, R, B, D, stream=s))
    for s, (kk, t) in tickets.items():
        ctx.wait(t, outs[kk])
    for o, e in zip(outs, exp):
        assert np.array_equal(o, e)
    assert np.array_equal(ctx.compute_sharded(frames[0][0], frames[0][1], B, D), exp[0])


def test_pinned_pool_zero_copy_path(ctx, oracle):
    rng = np.random.default_rng(10)
    L, R = synth_pair(rng, 100, 256, 1)
    pL = ctx.host_array((100, 256)); pR = ctx.host_array((100, 256)); pO = ctx.host_array((100, 256))
    pL[:] = L; pR[:] = R; pO[:] = 0
    ctx.compute(pL, pR, 7, 32, out=pO)
    assert np.array_equal(pO, oracle.frame_box(L, R, 7, 32))


def test_submit_into_pinned_destination(ctx, oracle):
    """sadgpu_submit_into: the D2H copy lands in the caller's pinned map, wait() copies nothing; row ranges assemble."""
    import despair
    rng = np.random.default_rng(11)
    frames = [synth_pair(rng, 96, 200, 1) for _ in range(6)]        # width 200: pitch == w (multiple of 4)
    pins = [ctx.host_pair(96, 200) for _ in frames]
    outs = [ctx.host_array((96, 200)) for _ in frames]
    for (pl, pr), (L, R), o in zip(pins, frames, outs):
        pl[:] = L; pr[:] = R; o[:] = 7
    tickets = {}
    for k in range(len(frames)):
        s = k % 4
        if s in tickets:
            ctx.wait(tickets.pop(s))
        tickets[s] = ctx.submit(pins[k][0], pins[k][1], 9, 64, stream=s, out=outs[k])
    for t in tickets.values():
        ctx.wait(t)
    for o, (L, R) in zip(outs, frames):
        assert np.array_equal(o, oracle.frame_box(L, R, 9, 64))
    # two disjoint row ranges of one map, odd width (2-D copies), rows outside stay untouched
    L, R = synth_pair(rng, 64, 101, 0)
    o = ctx.host_array((64, 101)); o[:] = 9
    t0 = ctx.submit(L, R, 5, 32, y0=0, y1=30, stream=0, out=o)
    t1 = ctx.submit(L, R, 5, 32, y0=30, y1=60, stream=1, out=o)
    ctx.wait(t0); ctx.wait(t1)
    exp = oracle.frame_box(L, R, 5, 32)
    assert np.array_equal(o[:60], exp[:60]) and (o[60:] == 9).all()
    with pytest.raises(despair.SadGpuError) as e:                   # pageable destination: refused, nothing is retained
        ctx.submit(L, R, 5, 32, stream=0, out=np.zeros((64, 101), np.uint8))
    assert e.value.code == -1
    t = ctx.submit(L, R, 5, 32, stream=0)
    with pytest.raises(despair.SadGpuError) as e:                   # plain submit needs a destination at wait
        ctx.wait(t)
    assert e.value.code == -1


@pytest.mark.parametrize("B,D", [(9, 128), (15, 128), (13, 256), (16, 64), (31, 128)])
def test_pipelined_kernels_are_deterministic_at_full_size(torch_mod, ctx, oracle, B, D):
    """The warp-specialised kernels hand rows between warps through barriers; a protocol slip shows up as a rare wrong
    tile, not as a crash.  Six launches of an 8-frame 1080p batch must agree bit for bit with each other, and a row range
    of the first and the last frame with the oracle."""
    rng = np.random.default_rng(80 + B)
    F, H, W = 8, 1080, 1920
    L = rng.integers(0, 256, (F, H, W), dtype=np.uint8); R = np.roll(L, -17, axis=2)
    R[:, ::3] = rng.integers(0, 256, (F, (H + 2) // 3, W), dtype=np.uint8)
    dL = torch_mod.from_numpy(L).cuda(); dR = torch_mod.from_numpy(R).cuda()
    st = torch_mod.cuda.current_stream().cuda_stream
    first = None
    for rep in range(6):
        dO = torch_mod.full((F, H, W), 77, dtype=torch_mod.uint8, device="cuda")
        ctx.compute_device_batch(F, dL.data_ptr(), W, W * H, dR.data_ptr(), W, W * H, W, H, B, D, dO.data_ptr(), W, W * H, cuda_stream=st)
        torch_mod.cuda.synchronize()
        if first is None:
            first = dO
            for f in (0, F - 1):
                assert np.array_equal(dO[f, 500:520].cpu().numpy(), oracle.frame_box(L[f], R[f], B, D, 500, 520)), (B, D, f)
        else:
            assert bool((dO == first).all()), (B, D, rep)


def test_submit_batch_video_stream(oracle):
    """sadgpu_submit_batch_into: n frame pairs per call, one DMA each way, one launch; pinned and pageable sources;
    chunked disparity range (key map per frame) and a wide window."""
    import despair
    c = despair.Context([0], 320, 120, 2)
    try:
        rng = np.random.default_rng(12)
        c.reserve_batch(4)                                               # stream buffers for 4 pairs; the 5-pair batch grows its stream
        for (B, D, n, pinned) in [(9, 128, 5, True), (15, 256, 4, False), (31, 40, 3, True), (5, 16, 1, False)]:
            frames = [synth_pair(rng, 120, 320, k % 3) for k in range(n)]
            pairs = c.host_array((n, 2, 120, 320)) if pinned else np.zeros((n, 2, 120, 320), np.uint8)
            for k, (L, R) in enumerate(frames):
                pairs[k, 0] = L; pairs[k, 1] = R
            out = c.host_array((n, 120, 320)); out[:] = 3
            c.wait(c.submit_batch(pairs, B, D, out, stream=n % 2))
            for k, (L, R) in enumerate(frames):
                assert np.array_equal(out[k], oracle.frame_box(L, R, B, D)), (B, D, k)
        with pytest.raises(despair.SadGpuError) as e:                   # pageable destination refused
            c.submit_batch(np.zeros((2, 2, 120, 320), np.uint8), 9, 64, np.zeros((2, 120, 320), np.uint8))
        assert e.value.code == -1
    finally:
        c.close()


def test_error_codes(ctx):
    import despair
    L = np.zeros((10, 10), np.uint8)
    for kw, code in [(dict(block_size=0, max_disparity=64), -1), (dict(block_size=33, max_disparity=64), -1),
                     (dict(block_size=9, max_disparity=0), -1), (dict(block_size=9, max_disparity=300), -1)]:
        with pytest.raises(despair.SadGpuError) as e:
            ctx.compute(L, L, **kw)
        assert e.value.code == code
    with pytest.raises(despair.SadGpuError) as e:
        ctx.compute(L, L, 9, 64, y0=4, y1=11)
    assert e.value.code == -2
    with pytest.raises(despair.SadGpuError) as e:
        ctx.compute(L, L, 9, 64, stream=99)
    assert e.value.code == -2
    big = np.zeros((3000, 10), np.uint8)
    with pytest.raises(despair.SadGpuError) as e:
        ctx.compute(big, big, 9, 64)
    assert e.value.code == -2
    with pytest.raises(despair.SadGpuError) as e:
        ctx.wait(12345 << 16, L)
    assert e.value.code == -4


def test_go_exact_gray_conversion_on_gpu(torch_mod, ctx):
    """the GPU luma kernel against the Go-exact formulas of oracle/go_image.py."""
    torch = torch_mod
    from oracle.go_image import _luma16
    rng = np.random.default_rng(12)
    for (h, w) in [(37, 101), (64, 256), (5, 3)]:
        rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        rgba[..., 3] = rng.choice([255, 254, 251, 128, 0, 1], size=(h, w))          # testdata alpha minimum is 251-254
        a = rgba.astype(np.uint64)
        c16 = lambda c: (c * 257 * a[..., 3]) // 255
        exp = _luma16(c16(a[..., 0]), c16(a[..., 1]), c16(a[..., 2]))
        d = torch.from_numpy(rgba).cuda(); g = torch.zeros((h, w), dtype=torch.uint8, device="cuda")
        ctx.gray_device(d.data_ptr(), w * 4, 4, 0, w, h, g.data_ptr(), w, cuda_stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(g.cpu().numpy(), exp)
        rgb = np.ascontiguousarray(rgba[..., :3]); b = rgb.astype(np.uint64)
        d3 = torch.from_numpy(rgb).cuda()
        ctx.gray_device(d3.data_ptr(), w * 3, 3, 1, w, h, g.data_ptr(), w, cuda_stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(g.cpu().numpy(), _luma16(b[..., 0] * 257, b[..., 1] * 257, b[..., 2] * 257))
        ctx.gray_device(d3.data_ptr(), w * 3, 3, 2, w, h, g.data_ptr(), w, cuda_stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert not g.cpu().numpy().any()                                               # gray.go:35-37 as written
    # every opaque colour (the folded-weights fast path) and every colour at three other alphas, 4096 x 4096 pixels each
    idx = np.arange(1 << 24, dtype=np.uint32)
    for alpha in (255, 254, 128, 3):
        rgba = np.empty((1 << 24, 4), np.uint8)
        rgba[:, 0] = idx & 255; rgba[:, 1] = (idx >> 8) & 255; rgba[:, 2] = idx >> 16; rgba[:, 3] = alpha
        a = rgba.astype(np.uint64)
        exp = _luma16((a[:, 0] * 257 * alpha) // 255, (a[:, 1] * 257 * alpha) // 255, (a[:, 2] * 257 * alpha) // 255).reshape(4096, 4096)
        d = torch.from_numpy(rgba.reshape(4096, 4096 * 4)).cuda(); g = torch.zeros((4096, 4096), dtype=torch.uint8, device="cuda")
        ctx.gray_device(d.data_ptr(), 4096 * 4, 4, 0, 4096, 4096, g.data_ptr(), 4096, cuda_stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(g.cpu().numpy(), exp), alpha


def test_compute_nrgba_colour_pair(ctx, oracle, manifest):
    """sadgpu_compute_nrgba: the RGBA crops cut from the reference testdata go in as image.NRGBA.Pix would,
    the luma is taken on the device, the map equals the oracle's on the Go-exact luma; pinned and pageable, odd strides."""
    from PIL import Image
    from oracle.go_image import load_png
    m = manifest["rgba_crop"]
    pl = os.path.join(GOLDEN, f"L_{m['tag']}_rgba_crop.png"); pr = os.path.join(GOLDEN, f"R_{m['tag']}_rgba_crop.png")
    L4 = np.array(Image.open(pl)); R4 = np.array(Image.open(pr))
    assert L4.shape == (m["h"], m["w"], 4) and int((L4[..., 3] < 255).sum() + (R4[..., 3] < 255).sum()) == m["pixels_with_alpha_below_255"]
    gl = load_png(pl, "intended"); gr = load_png(pr, "intended")
    for (B, D) in ((9, 64), (16, 64), (15, 128)):
        exp = oracle.frame_box(gl, gr, B, D)
        assert np.array_equal(ctx.compute_nrgba(L4, R4, B, D), exp), (B, D)
    assert sha(ctx.compute_nrgba(L4, R4, 9, 64)) == m["b9_d64_sha256"]
    # row stride larger than 4*w (a sub-image of a wider NRGBA), odd width, pinned destination
    wide = np.zeros((m["h"], m["w"] + 7, 4), np.uint8); wide[:, 3:3 + 301] = L4[:, :301]
    wide_r = np.zeros_like(wide); wide_r[:, 3:3 + 301] = R4[:, :301]
    out = ctx.host_array((m["h"], 301))
    ctx.compute_nrgba(wide[:, 3:3 + 301], wide_r[:, 3:3 + 301], 9, 64, out=out)
    assert np.array_equal(out, oracle.frame_box(np.ascontiguousarray(gl[:, :301]), np.ascontiguousarray(gr[:, :301]), 9, 64))


def test_row_band_sharding_across_devices(torch_mod, oracle):
    """cfg4-style: one frame split into row bands with a block-size halo over every visible GPU of the box
    (sadgpu_compute_sharded), host-side gather; frames of a stream round-robin over devices (stream s -> device s % n)."""
    import despair
    n = torch_mod.cuda.device_count()
    devs = list(range(min(n, 8)))
    c = despair.Context(devs, 640, 400, max(2, len(devs)))
    rng = np.random.default_rng(21)
    L, R = synth_pair(rng, 400, 640, 1)
    exp = oracle.frame_box(L, R, 31, 64)
    assert np.array_equal(c.compute_sharded(L, R, 31, 64), exp)
    outs = [np.zeros_like(L) for _ in devs]
    ts = [c.submit(L, R, 9, 128, stream=s) for s in range(len(devs))]        # stream s lives on device s % n
    for s, t in enumerate(ts):
        c.wait(t, outs[s])
    exp2 = oracle.frame_box(L, R, 9, 128)
    assert all(np.array_equal(o, exp2) for o in outs)
    c.close()


@pytest.mark.parametrize("variant,B,D", [(2, 13, 64), (2, 9, 16), (3, 9, 128), (3, 5, 200), (3, 9, 64), (3, 7, 32), (3, 3, 16), (3, 8, 48),
                                         (3, 11, 64), (3, 13, 128), (3, 15, 256), (3, 16, 64), (3, 17, 200), (3, 14, 16), (3, 11, 32), (3, 15, 20),
                                         (4, 31, 256), (4, 16, 33),
                                         (6, 15, 256), (6, 11, 128), (6, 13, 40), (6, 31, 256), (6, 16, 64), (6, 22, 100),
                                         (7, 31, 256), (7, 19, 32), (7, 16, 16)])
def test_unaligned_pitch_and_odd_widths(torch_mod, ctx, oracle, variant, B, D):
    """Pitches that are not multiples of 4 disable the aligned 32-bit tile loads; widths that are not multiples of the
    strip width exercise the right-edge masking (sad.go:231-233) in every kernel."""
    rng = np.random.default_rng(70 + variant + B)
    for (H, W, pad) in [(33, 131, 3), (20, 257, 1), (41, 64, 5), (17, 95, 2)]:
        L, R = synth_pair(rng, H, W, 1)
        got = dev_run(torch_mod, ctx, L, R, B, D, dict(kernel_variant=variant), pitch_pad=pad)
        assert np.array_equal(got, oracle.frame_box(L, R, B, D)), (H, W, pad, variant, B, D)


def test_tma_tile_loader_matches_plain_loader(torch_mod, ctx, oracle):
    """The warp-specialised kernel loads its tiles with TMA (cp.async.bulk.tensor, hardware zero fill) when base, pitch and
    frame stride are 16-byte aligned, otherwise with plain loads; both must agree with the oracle at image borders."""
    import ctypes
    from despair import _native as N
    torch = torch_mod
    rng = np.random.default_rng(33)
    st = torch.cuda.current_stream().cuda_stream
    for (W, H, B, D, F) in [(64, 20, 9, 128, 1), (128, 37, 9, 128, 2), (320, 50, 5, 200, 3), (96, 9, 1, 68, 1), (160, 33, 8, 256, 2), (48, 70, 9, 100, 1),
                            (640, 48, 9, 64, 2), (208, 31, 7, 40, 1), (400, 25, 9, 32, 2), (112, 40, 5, 16, 3), (336, 19, 3, 8, 1),
                            (256, 60, 11, 64, 2), (160, 45, 15, 128, 1), (192, 50, 16, 64, 2), (96, 70, 17, 256, 1), (128, 33, 13, 20, 1),
                            (320, 41, 15, 32, 2), (64, 30, 10, 16, 1)]:
        Ls = rng.integers(0, 256, (F, H, W), dtype=np.uint8); Rs = rng.integers(0, 256, (F, H, W), dtype=np.uint8)
        dL = torch.from_numpy(Ls).cuda(); dR = torch.from_numpy(Rs).cuda()
        for no_tma in (0, 1):
            dO = torch.zeros_like(dL)
            t = N.Tuning(); t.kernel_variant = 3; t.reserved[2] = no_tma
            N.check(N.lib().sadgpu_compute_device_batch(ctx._h, 0, F, dL.data_ptr(), W, W * H, dR.data_ptr(), W, W * H, W, H, B, D,
                                                        0, H, dO.data_ptr(), W, W * H, st, ctypes.byref(t)))
            torch.cuda.synchronize()
            got = dO.cpu().numpy()
            for f in range(F):
                assert np.array_equal(got[f], oracle.frame_box(Ls[f], Rs[f], B, D)), (W, H, B, D, f, no_tma)


UI_BLOCKS = list(range(3, 32, 2)) + [16]                 # cmd/components/control.templ:25-27 (odd 3..31) + the start-up default 16
UI_DISPARITIES = list(range(16, 257, 16))                # control.templ:71-73


@pytest.mark.parametrize("B", UI_BLOCKS)
def test_full_ui_grid_with_planner_defaults(torch_mod, ctx, oracle, B):
    """Every point of the WebUI's parameter grid through the planner's own choice of kernel, chunking and tiling (no tuning):
    the variant / chunk switch points (D 64|68, 128|132, 200|204, 17/18/33-group slots) are all crossed.  Inputs rotate through
    random, shifted texture, tie-heavy, 255-vs-0 and flat images."""
    rng = np.random.default_rng(900 + B)
    for i, D in enumerate(UI_DISPARITIES):
        W = int(rng.integers(180, 230)); H = int(rng.integers(56, 72))
        L, R = synth_pair(rng, H, W, (i + B) % 5)
        assert np.array_equal(dev_run(torch_mod, ctx, L, R, B, D), oracle.frame_box(L, R, B, D)), (W, H, B, D)


@pytest.mark.parametrize("B", UI_BLOCKS)
def test_full_ui_grid_through_the_host_entry_point(ctx, oracle, B):
    """The same grid through sadgpu_compute (host buffers, odd widths): inside the library every plane has a 16-byte pitch, so the
    TMA kernels — the planner's first choice at every point — run whatever the width."""
    rng = np.random.default_rng(1900 + B)
    for i, D in enumerate(UI_DISPARITIES):
        W = int(rng.integers(150, 260)); H = int(rng.integers(40, 80))
        L, R = synth_pair(rng, H, W, (i + B + 2) % 5)
        assert np.array_equal(ctx.compute(L, R, B, D), oracle.frame_box(L, R, B, D)), (W, H, B, D)


def _bands(h, chunk):
    return [(y, min(y + chunk, h)) for y in range(0, h, chunk)]


def test_compute_region_chunks_share_one_pass(ctx, oracle):
    """sadgpu_compute_region: the reference's chunkings (OutputCamera row bands, output.go:172-187; RunSad tiles, sad.go:128-153)
    through region-local buffers; all chunks of a frame pair cost ONE whole-frame GPU pass."""
    rng = np.random.default_rng(31)
    for (H, W, B, D) in [(480, 640, 16, 64), (270, 481, 9, 128), (96, 130, 31, 256)]:
        L, R = synth_pair(rng, H, W, 1)
        exp = oracle.frame_box(L, R, B, D)
        s0 = ctx.region_stats()
        out = np.zeros_like(L)
        for (x0, y0, x1, y1) in oracle.output_camera_chunks(W, H, 32):
            blk = ctx.compute_region(L, R, B, D, (x0, y0, x1, y1))
            assert blk.shape == (y1 - y0, x1 - x0)
            out[y0:y1, x0:x1] = blk
        assert np.array_equal(out, exp)
        s1 = ctx.region_stats()
        assert s1["frames"] - s0["frames"] == 1 and s1["calls"] - s0["calls"] == len(oracle.output_camera_chunks(W, H, 32))
        out2 = np.zeros_like(L)
        for (x0, y0, x1, y1) in oracle.run_sad_chunks(W, H, 4):
            out2[y0:y1, x0:x1] = ctx.compute_region(L, R, B, D, (x0, y0, x1, y1))
        assert np.array_equal(out2, exp)
        assert ctx.region_stats()["frames"] - s1["frames"] == 1      # the first sweep served every pixel: its entry was retired
    # destination with a row stride larger than the region (a window of a bigger buffer), strided sources
    big = np.zeros((40, 300), np.uint8)
    Lw = np.zeros((96, 200), np.uint8); Rw = np.zeros((96, 200), np.uint8)
    L, R = synth_pair(rng, 96, 130, 0)
    Lw[:, 7:137] = L; Rw[:, 9:139] = R
    ctx.compute_region(Lw[:, 7:137], Rw[:, 9:139], 7, 40, (11, 20, 121, 60), out=big[:, 50:160])
    assert np.array_equal(big[:, 50:160], oracle.frame_box(L, R, 7, 40)[20:60, 11:121]) and not big[:, :50].any() and not big[:, 160:].any()


def test_compute_region_concurrent_workers(ctx, oracle):
    """32 worker threads (pkg/camera/output.go:37) pull the 160 bands of a 480-row frame, two frame pairs interleaved."""
    import threading, queue
    rng = np.random.default_rng(32)
    H, W, B, D = 480, 640, 16, 64
    pairs = [synth_pair(rng, H, W, 1) for _ in range(2)]
    exps = [oracle.frame_box(L, R, B, D) for L, R in pairs]
    outs = [np.zeros((H, W), np.uint8) for _ in pairs]
    q = queue.Queue()
    for (y0, y1) in _bands(H, max(1, H // 128)):
        for k in range(2):
            q.put((k, y0, y1))
    errors = []
    s0 = ctx.region_stats()

    def worker():
        try:
            while True:
                try:
                    k, y0, y1 = q.get_nowait()
                except queue.Empty:
                    return
                outs[k][y0:y1] = ctx.compute_region(pairs[k][0], pairs[k][1], B, D, (0, y0, W, y1))
        except Exception as e:      # noqa: BLE001
            errors.append(e)

    ts = [threading.Thread(target=worker) for _ in range(32)]
    [t.start() for t in ts]; [t.join() for t in ts]
    assert not errors, errors
    for o, e in zip(outs, exps):
        assert np.array_equal(o, e)
    s1 = ctx.region_stats()
    assert s1["calls"] - s0["calls"] == 320 and s1["frames"] - s0["frames"] == 2 and s1["stale"] == s0["stale"]


def test_compute_region_never_serves_a_stale_frame(ctx, oracle):
    """A caller that refills the SAME image objects with new pixels between chunks (or between frames) gets the new frame:
    every chunk compares the rows it depends on with the snapshot before it is served from a shared pass."""
    rng = np.random.default_rng(33)
    H, W, B, D = 120, 200, 9, 64
    L, R = synth_pair(rng, H, W, 1)
    L2, R2 = synth_pair(rng, H, W, 1)
    bands = _bands(H, 10)
    s0 = ctx.region_stats()
    out = np.zeros((H, W), np.uint8)
    for (y0, y1) in bands[:5]:
        out[y0:y1] = ctx.compute_region(L, R, B, D, (0, y0, W, y1))
    exp_old = oracle.frame_box(L, R, B, D)
    assert np.array_equal(out[:50], exp_old[:50])
    L[:] = L2; R[:] = R2                                   # same objects, same addresses, new pixels
    exp_new = oracle.frame_box(L, R, B, D)
    for (y0, y1) in bands[5:]:
        out[y0:y1] = ctx.compute_region(L, R, B, D, (0, y0, W, y1))
    assert np.array_equal(out[50:], exp_new[50:])
    s1 = ctx.region_stats()
    assert s1["stale"] - s0["stale"] >= 1 and s1["frames"] - s0["frames"] == 2
    # a change OUTSIDE the rows a chunk depends on does not invalidate it; a change inside the halo does
    out[:] = 0
    out[0:10] = ctx.compute_region(L, R, B, D, (0, 0, W, 10))
    L[100, 5] ^= 0xFF
    assert np.array_equal(ctx.compute_region(L, R, B, D, (0, 10, W, 20)), exp_new[10:20])
    assert ctx.region_stats()["stale"] == s1["stale"]
    L[23, 17] ^= 0x55                                      # row 23 is inside the h = 4 halo of rows [10, 20)
    assert np.array_equal(ctx.compute_region(L, R, B, D, (0, 10, W, 20)), oracle.frame_box(L, R, B, D)[10:20])
    assert ctx.region_stats()["stale"] == s1["stale"] + 1


def test_compute_region_errors(ctx):
    import despair
    L = np.zeros((20, 30), np.uint8)
    for region, B, D, code in [((0, 0, 30, 20), 33, 64, -1), ((0, 0, 30, 20), 9, 0, -1), ((0, 0, 31, 20), 9, 64, -2),
                               ((0, 5, 30, 21), 9, 64, -2), ((4, 0, 2, 20), 9, 64, -2)]:
        with pytest.raises(despair.SadGpuError) as e:
            ctx.compute_region(L, L, B, D, region)
        assert e.value.code == code, (region, B, D)
    assert ctx.compute_region(L, L, 9, 64, (5, 5, 5, 10)).shape == (5, 0)          # empty region: nothing to do


def test_concurrent_device_calls_with_chunked_disparity_range(torch_mod, ctx, oracle):
    """Two host threads, two CUDA streams, ONE device, max_disparity 256 (two chunks merged through a key map): each call has
    its own stream-ordered scratch, so concurrent calls cannot corrupt each other (round-1 advisor finding)."""
    import threading
    torch = torch_mod
    rng = np.random.default_rng(34)
    H, W, F = 200, 640, 3
    jobs = []
    for (B, D) in [(9, 256), (15, 256)]:
        L = rng.integers(0, 256, (F, H, W), dtype=np.uint8); R = np.roll(L, -23, axis=2).copy()
        R[:, ::5] = rng.integers(0, 256, (F, (H + 4) // 5, W), dtype=np.uint8)
        exp = np.stack([oracle.frame_box(L[f], R[f], B, D) for f in range(F)])
        jobs.append((B, D, torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda(), exp))
    torch.cuda.synchronize()
    errors = []

    def run(job):
        B, D, dL, dR, exp = job
        try:
            st = torch.cuda.Stream()
            for rep in range(12):
                dO = torch.full((F, H, W), 99, dtype=torch.uint8, device="cuda")
                st.wait_stream(torch.cuda.current_stream())
                ctx.compute_device_batch(F, dL.data_ptr(), W, W * H, dR.data_ptr(), W, W * H, W, H, B, D, dO.data_ptr(), W, W * H,
                                         cuda_stream=st.cuda_stream)
                st.synchronize()
                if not np.array_equal(dO.cpu().numpy(), exp):
                    errors.append((B, D, rep))
        except Exception as e:      # noqa: BLE001
            errors.append(e)

    ts = [threading.Thread(target=run, args=(j,)) for j in jobs]
    [t.start() for t in ts]; [t.join() for t in ts]
    assert not errors, errors


def test_wait_uploaded_releases_pinned_inputs(ctx, oracle):
    """Pinned-pool inputs are borrowed by the upload: after sadgpu_wait_uploaded the caller may refill them while the frame
    is still on the GPU (one pinned frame per camera)."""
    rng = np.random.default_rng(35)
    H, W = 200, 320
    pl, pr = ctx.host_pair(H, W); out = ctx.host_array((H, W))
    frames = [synth_pair(rng, H, W, 1) for _ in range(4)]
    for (L, R) in frames:
        pl[:] = L; pr[:] = R
        t = ctx.submit(pl, pr, 9, 128, stream=0, out=out)
        ctx.wait_uploaded(t)
        pl[:] = 0; pr[:] = 255                              # the next capture overwrites the frame
        ctx.wait(t)
        assert np.array_equal(out, oracle.frame_box(L, R, 9, 128))


def test_post_processing_hooks_are_additive(torch_mod, ctx, oracle):
    """3x3 median and left-right consistency on the device against their numpy statement (oracle/post_numpy.py);
    the hooks leave the bit-exact entry points untouched."""
    from oracle import post_numpy as P
    torch = torch_mod
    rng = np.random.default_rng(50)
    st = torch.cuda.current_stream().cuda_stream
    for (h, w) in [(37, 101), (64, 256), (5, 3), (1, 1)]:
        m = rng.integers(0, 256, (h, w), dtype=np.uint8); m2 = rng.integers(0, 256, (h, w), dtype=np.uint8)
        d = torch.from_numpy(m).cuda(); d2 = torch.from_numpy(m2).cuda(); o = torch.zeros_like(d)
        ctx.median3_device(d.data_ptr(), w, w, h, o.data_ptr(), w, cuda_stream=st)
        torch.cuda.synchronize()
        assert np.array_equal(o.cpu().numpy(), P.median3(m)), (h, w)
        for (D, tol, inv) in ((64, 1, 0), (256, 0, 255), (16, 2, 7)):
            ctx.lrcheck_device(d.data_ptr(), w, d2.data_ptr(), w, w, h, D, tol, inv, o.data_ptr(), w, cuda_stream=st)
            torch.cuda.synchronize()
            assert np.array_equal(o.cpu().numpy(), P.lr_check(m, m2, D, tol, inv)), (h, w, D, tol)
    # the host call: both maps through the bit-exact path, check, median
    base = rng.integers(0, 256, (90, 260), dtype=np.uint8)
    L = np.ascontiguousarray(base[:, 40:240]); R = np.ascontiguousarray(np.roll(base, -12, 1)[:, 40:240])
    for (B, D, med) in ((9, 64, False), (16, 64, True), (5, 32, True)):
        lm = oracle.frame_box(L, R, B, D)
        rm = P.right_referenced(oracle.frame_box, L, R, B, D)
        exp = P.lr_check(lm, rm, D, 1, 0)
        if med:
            exp = P.median3(exp)
        assert np.array_equal(ctx.compute_checked(L, R, B, D, tolerance=1, invalid_value=0, median=med), exp), (B, D, med)
        assert np.array_equal(ctx.compute(L, R, B, D), lm)                     # the plain path is what it was
    # on a clean shifted pair the check keeps the interior and removes the left band where the match leaves the image
    got = ctx.compute_checked(L, R, 9, 64, tolerance=1, invalid_value=0)
    d12 = 12 * 255 // 64
    assert (got[10:-10, 90:190] == d12).mean() > 0.95


def test_devices_from_environment(torch_mod, oracle, monkeypatch):
    """sadgpu_create(devices = NULL): ordinals 0..n-1, or the first n entries of SADGPU_DEVICES (how the C++ mirror and the Go stub
    pick their GPUs); a list that is too short or names a device that does not exist is an error, never a silent default."""
    import ctypes
    from despair import _native as N
    L = N.lib()
    rng = np.random.default_rng(8)
    Limg, Rimg = synth_pair(rng, 40, 96, 1)
    exp = oracle.frame_box(Limg, Rimg, 9, 64)

    def create(n):
        h = ctypes.c_void_p()
        rc = L.sadgpu_create(None, n, 128, 64, 1, ctypes.byref(h))
        return rc, h

    last = torch_mod.cuda.device_count() - 1
    for env in (None, "0", f"{last}", f" {last},0"):
        if env is None:
            monkeypatch.delenv("SADGPU_DEVICES", raising=False)
        else:
            monkeypatch.setenv("SADGPU_DEVICES", env)
        rc, h = create(1)
        assert rc == 0, (env, rc)
        out = np.zeros_like(Limg)
        N.check(L.sadgpu_compute(h, 0, Limg.ctypes.data, 96, Rimg.ctypes.data, 96, 96, 40, 9, 64, 0, 40, out.ctypes.data, 96))
        L.sadgpu_destroy(h)
        assert np.array_equal(out, exp), env
    for env, n in (("0", 2), ("99", 1), ("-1", 1), ("x", 1)):
        monkeypatch.setenv("SADGPU_DEVICES", env)
        rc, h = create(n)
        assert rc != 0, (env, n)


def test_concurrent_pageable_calls_share_the_staging_helpers(ctx, oracle):
    """Four threads, each on its own stream slot, call sadgpu_compute on pageable 1080p planes at the same time: the staging copies
    (>= 512 KB) are shared with whichever helper threads are idle at that moment, a helper another caller holds is not used."""
    import threading
    rng = np.random.default_rng(44)
    H, W, B, D = 1080, 1920, 9, 128
    pairs = [synth_pair(rng, H, W, 1) for _ in range(4)]
    exps = [oracle.frame_box(L, R, B, D, 500, 516) for L, R in pairs]
    errors = []

    def worker(k):
        try:
            out = np.zeros((H, W), np.uint8)
            for it in range(6):
                L, R = pairs[(k + it) % 4]
                ctx.compute(L, R, B, D, stream=k, out=out)
                if not np.array_equal(out[500:516], exps[(k + it) % 4]):
                    errors.append((k, it))
        except Exception as e:      # noqa: BLE001
            errors.append(e)

    ts = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    [t.start() for t in ts]; [t.join() for t in ts]
    assert not errors, errors
