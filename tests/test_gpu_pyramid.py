"""GPU tier: K1 (pyrDown + Scharr, TMA-staged) must be BIT-EXACT with the oracle / cv2 fixtures.  All calls go
through the C ABI (mcmtt_opticalflow_b200.klt -> libpsnw_klt.so)."""
import numpy as np
import pytest

from mcmtt_opticalflow_b200 import klt, synth

pytestmark = pytest.mark.gpu


def _check_levels(ctx, oracle, img, cam=0, slot=0, apron=26):
    cur = img
    for l in range(ctx.levels):
        if l:
            cur = oracle.pyr_down(cur)
        g_img, g_dxy = ctx.read_level(cam, slot, l)
        assert g_img.shape == cur.shape
        assert np.array_equal(g_img, cur), "u8 level %d differs" % l
        assert np.array_equal(g_dxy, oracle.scharr(cur)), "Scharr level %d differs" % l
        _check_apron(ctx, cur, cam, slot, l, apron)


def _check_apron(ctx, img, cam, slot, level, apron=26):
    """Levels >= 1 keep a REFLECT_101 apron in HBM (what copyMakeBorder builds around every level inside
    calcOpticalFlowPyrLK); every apron pixel a single reflection defines must equal np.pad(mode='reflect')."""
    pad, a = ctx.read_level_apron(cam, slot, level)
    h, w = img.shape
    assert a == (apron if level else 0)
    assert np.array_equal(pad[a:a + h, a:a + w], img)
    if not a:
        return
    ax, ay = min(a, w - 1), min(a, h - 1)               # reflection depth available on this level
    want = np.pad(img, ((ay, ay), (ax, ax)), mode="reflect")
    got = pad[a - ay:a + h + ay, a - ax:a + w + ax]
    assert np.array_equal(got, want), "apron of level %d differs" % level


@pytest.mark.parametrize("mode", ["level", "fused", "hybrid", "walk"])
@pytest.mark.parametrize("w,h,levels", [(768, 576, 4), (767, 575, 4), (129, 33, 3), (16, 16, 2), (130, 70, 3),
                                        (1920, 1080, 4), (257, 131, 5), (640, 480, 6), (193, 65, 4), (385, 129, 3),
                                        (3840, 2160, 5)])
def test_bit_exact_vs_oracle(oracle, monkeypatch, w, h, levels, mode):
    """Both pyramid builders -- one launch per level (pyramid.cu) and the fused multi-level kernel (pyramid_fused.cu) --
    are forced in turn (PSNW_PYR_MODE is read at psnw_create); the default picks by launch size."""
    monkeypatch.setenv("PSNW_PYR_MODE", mode)
    rng = np.random.default_rng(w * 1000 + h)
    img = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
    with klt.KLTContext(w, h, n_cameras=1, levels=levels) as ctx:
        ctx.push_frame(img)
        _check_levels(ctx, oracle, img)


@pytest.mark.parametrize("mode", ["level", "fused", "hybrid", "walk"])
@pytest.mark.parametrize("apron,w,h,levels", [(64, 768, 576, 4), (40, 333, 97, 4), (160, 640, 480, 5), (5, 130, 70, 3)])
def test_apron_width_is_a_context_parameter(oracle, monkeypatch, apron, w, h, levels, mode):
    """psnw_params.apron: wider aprons for contexts that track large windows (clamped to [26, 160])."""
    monkeypatch.setenv("PSNW_PYR_MODE", mode)
    rng = np.random.default_rng(apron * 7 + w)
    img = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
    with klt.KLTContext(w, h, n_cameras=1, levels=levels, apron=apron) as ctx:
        ctx.push_frame(img)
        _check_levels(ctx, oracle, img, apron=min(max(apron, 26), 160))


def test_golden_fixture(golden):
    a = golden["a"]
    with klt.KLTContext(a.shape[1], a.shape[0], levels=4) as ctx:
        ctx.push_frame(a)
        for l in range(4):
            img, dxy = ctx.read_level(0, 0, l)
            assert np.array_equal(img, golden["pyr_a_%d" % l])
            assert np.array_equal(dxy[..., 0], golden["dx_a_%d" % l])
            assert np.array_equal(dxy[..., 1], golden["dy_a_%d" % l])
    o = golden["pyr_odd_0"]
    with klt.KLTContext(o.shape[1], o.shape[0], levels=4) as ctx:
        ctx.push_frame(o)
        for l in range(4):
            img, dxy = ctx.read_level(0, 0, l)
            assert np.array_equal(img, golden["pyr_odd_%d" % l])
            assert np.array_equal(dxy[..., 0], golden["dx_odd_%d" % l])
            assert np.array_equal(dxy[..., 1], golden["dy_odd_%d" % l])


@pytest.mark.parametrize("mode", ["level", "fused", "hybrid", "walk"])
def test_batched_cameras_and_ring_rotation(oracle, monkeypatch, mode):
    monkeypatch.setenv("PSNW_PYR_MODE", mode)
    w, h, ncam = 384, 288, 3
    cams = [synth.Camera(w, h, 4, cam=c) for c in range(ncam)]
    with klt.KLTContext(w, h, n_cameras=ncam, levels=3) as ctx:
        frames = {}
        for t in range(6):                                   # ring of 4: frames 2..5 survive
            fr = [c.frame(t) for c in cams]
            frames[t] = fr
            ctx.push_frames(fr, frame_idx=t)
            assert ctx.frames_held(0) == min(t + 1, 4)
        for c in range(ncam):
            for slot in range(4):                            # slot k == k frames ago (Tracker2D.cpp:311-316)
                _check_levels(ctx, oracle, frames[5 - slot][c], cam=c, slot=slot)
        # pushing a subset of cameras only rotates those
        ctx.push_frames([frames[0][1]], cams=[1])
        _check_levels(ctx, oracle, frames[0][1], cam=1, slot=0)
        _check_levels(ctx, oracle, frames[5][1], cam=1, slot=1)
        _check_levels(ctx, oracle, frames[5][0], cam=0, slot=0)


def test_pitched_host_input(oracle):
    w, h = 200, 100
    big = np.random.default_rng(1).integers(0, 256, size=(h, 512), dtype=np.uint8)
    view = big[:, :w]
    with klt.KLTContext(w, h, levels=2) as ctx:
        import ctypes as C
        cams = np.zeros(1, np.int32)
        ptrs = (C.c_void_p * 1)(view.ctypes.data)
        assert ctx.lib.psnw_push_frames(ctx._h, 1, cams.ctypes.data, ptrs, 512, 0, klt.MEM_HOST) == 0
        _check_levels(ctx, oracle, np.ascontiguousarray(view))


@pytest.mark.parametrize("bits", [14, 15])
def test_bgr_intake(oracle, bits):
    w, h = 321, 123
    bgr = np.random.default_rng(2).integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    b, g, r = [bgr[..., i].astype(np.int64) for i in range(3)]
    if bits == 15:
        gray = (3735 * b + 19235 * g + 9798 * r + (1 << 14)) >> 15
    else:
        gray = (1868 * b + 9617 * g + 4899 * r + (1 << 13)) >> 14
    gray = gray.astype(np.uint8)
    with klt.KLTContext(w, h, levels=3) as ctx:
        ctx.push_frames_bgr([bgr], luma_bits=bits)
        _check_levels(ctx, oracle, gray)


def test_errors():
    with klt.KLTContext(64, 64, levels=2) as ctx:
        with pytest.raises(klt.PsnwError) as e:
            ctx.read_level(0, 0, 0)                          # nothing pushed yet
        assert e.value.code == klt.ESTATE
        with pytest.raises(klt.PsnwError) as e:
            ctx.push_frames([np.zeros((64, 64), np.uint8)] * 2, cams=[0, 0])
        assert e.value.code == klt.EARG
    with pytest.raises(klt.PsnwError):
        klt.KLTContext(4, 4)


def test_contiguous_block_takes_the_single_copy_path(oracle):
    """[cameras][h][w] block with w == pitch: psnw_push_frames moves all frames with one strided copy."""
    w, h, ncam = 256, 96, 5
    blk = np.random.default_rng(3).integers(0, 256, size=(ncam, h, w), dtype=np.uint8)
    with klt.KLTContext(w, h, n_cameras=ncam, levels=3) as ctx:
        for rep in range(2):                                   # twice: the ring advances in lockstep
            ctx.push_frames([blk[c] for c in range(ncam)])
            for c in range(ncam):
                _check_levels(ctx, oracle, blk[c], cam=c, slot=0)
        ctx.push_frames([blk[c] for c in (4, 2, 0)], cams=[4, 2, 0])     # negative source stride, non-unit camera step
        for c in (4, 2, 0):
            _check_levels(ctx, oracle, blk[c], cam=c, slot=0)


@pytest.mark.parametrize("ctas", ["0", "1"])
def test_level_walker_many_frames(oracle, monkeypatch, ctas):
    """The persistent level walker (one launch for every level of every frame, levels chained through per-frame completion
    counters): a batch with far more tiles than resident CTAs, pushed repeatedly with changing content so that a tile read
    before its producers finished would show as stale pixels; one CTA per SM forces long waits on the counters."""
    monkeypatch.setenv("PSNW_PYR_MODE", "walk")
    monkeypatch.setenv("PSNW_WALK_CTAS", ctas)
    w, h, ncam = 333, 201, 40
    rng = np.random.default_rng(11)
    with klt.KLTContext(w, h, n_cameras=ncam, levels=4, ring_slots=2) as ctx:
        for rep in range(3):
            blk = rng.integers(0, 256, size=(ncam, h, w), dtype=np.uint8)
            ctx.push_frames([blk[c] for c in range(ncam)])
            for c in range(0, ncam, 3):
                _check_levels(ctx, oracle, blk[c], cam=c, slot=0)


def test_consecutive_builds_of_other_planes_overlap(oracle, monkeypatch):
    """Back-to-back fused builds of disjoint ring slots start without waiting for each other (late_wait) and complete in
    stream order; a build that revisits planes of the chain waits first.  Results are those of the serial order."""
    monkeypatch.setenv("PSNW_PYR_MODE", "fused")
    w, h, ncam, slots = 768, 576, 4, 4        # (large enough for the taller 256-thread regions chained builds switch to)
    rng = np.random.default_rng(5)
    frames = [rng.integers(0, 256, size=(ncam, h, w), dtype=np.uint8) for _ in range(slots)]
    with klt.KLTContext(w, h, n_cameras=ncam, levels=4, ring_slots=slots) as ctx:
        for s in range(slots):
            ctx.push_frames([frames[s][c] for c in range(ncam)])
        cams = np.arange(ncam, dtype=np.int32)
        for rep in range(3):
            for s in range(slots):
                ctx.rebuild_pyramids(cams, slot=s)
            ctx.rebuild_pyramids(cams, slot=0)                    # planes already in the chain: waits first
            ctx.rebuild_pyramids(cams[:1], slot=1)
            ctx.rebuild_pyramids(cams[1:], slot=1)                # other cameras of the same slot: disjoint planes
        ctx.sync()
        for s in range(slots):
            for c in range(ncam):
                _check_levels(ctx, oracle, frames[slots - 1 - s][c], cam=c, slot=s)


def test_frames_in_separate_pinned_allocations(oracle):
    """Two frames are always 'at equal distances' in memory; when they live in different page-locked allocations the single
    strided copy is refused by the runtime and every frame takes its own copy (found by the stream-ordering fuzz test)."""
    import torch
    w, h = 256, 192                                            # w == pitch: the single-copy path is considered
    rng = np.random.default_rng(17)
    bufs = [torch.from_numpy(rng.integers(0, 256, size=(h, w), dtype=np.uint8)).pin_memory() for _ in range(6)]
    with klt.KLTContext(w, h, n_cameras=3, levels=3) as ctx:
        for order in ([0, 3, 5], [5, 1, 2], [4, 4, 0]):        # ascending, mixed, the same source frame for two cameras
            ctx.push_frames([bufs[i].data_ptr() for i in order], [0, 1, 2], 0, klt.MEM_PINNED, pitch=w)
            ctx.sync()
            for c, i in enumerate(order):
                _check_levels(ctx, oracle, bufs[i].numpy(), cam=c, slot=0)
