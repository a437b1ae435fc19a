"""GPU tier: the BASELINE.json configurations as RIGS -- C1 (768x576, 20 targets x 100 seeds, 3 levels), C3 (1920x1080,
200 targets x 256 seeds, 4 levels) and C4 (3840x2160, stride-4 grid, 5 levels; a >= 50 000-seed subsample so that the CPU
checkers finish in seconds) -- through the C ABI, compared with
  (a) the exact-sum oracle (oracle/klt_oracle.c, the kernel's own arithmetic): identical status, <= 1e-4 px;
  (b) live cv2 4.13 (the runnable build of the library the reference links at Tracker2D.cpp:776-782, :871-877) when it is
      importable on the box: north_star's gates -- status agreement >= 99.5 %, <= 0.05 px for >= 99 % of agreeing points,
      per-target box IoU >= 0.99 (boxes from LocalSearchKLT on each side's own tracks).
C2 is covered by tests/test_gpu_lk.py."""
import numpy as np
import pytest

from mcmtt_opticalflow_b200 import klt, synth

pytestmark = pytest.mark.gpu

TOL_PX, MIN_STATUS_AGREE, MIN_WITHIN_TOL, MIN_IOU = 0.05, 0.995, 0.99, 0.99


def _compare(ctx, oracle, f0, f1, pts, levels, boxes=None, off=None, cam=0):
    o, st, er, fb = ctx.track(pts, cam=cam)
    ml = levels - 1
    pa, pb = oracle.build_pyramid(f0, ml), oracle.build_pyramid(f1, ml)
    g_o, g_st, g_er, g_fb = oracle.track_fb(pa, pb, pts, (21, 21), ml, fb_thresh=1.0, exact_sums=True)
    assert np.array_equal(st, g_st), int((st != g_st).sum())
    assert np.linalg.norm(o - g_o, axis=1).max() <= 1e-4
    assert np.abs(er - g_er).max() <= 1e-5 and np.abs(fb - g_fb).max() <= 1e-4
    nb = None
    if boxes is not None:
        nb, mask, cnt = ctx.box_update(boxes, pts, o, off, status=st)
        for t in range(len(boxes)):
            s = slice(off[t], off[t + 1])
            b, inl = oracle.local_search_klt(boxes[t], pts[s], g_o[s], g_st[s])
            assert np.array_equal(nb[t], b) and np.array_equal(np.flatnonzero(mask[s]), inl), t
    stats = {"tracked": float(st.mean())}
    try:
        from oracle import cv2_ref
    except Exception:                     # cv2 missing on the box: the oracle comparison above stands alone
        return stats
    c_o, c_st, c_er, c_fb = cv2_ref.track_fb(f0, f1, pts, (21, 21), ml, fb_thresh=1.0)
    agree = float((st == c_st).mean())
    ok = (st == 1) & (c_st == 1)
    d = np.linalg.norm(o - c_o, axis=1)[ok]
    assert agree >= MIN_STATUS_AGREE, agree
    assert (d <= TOL_PX).mean() >= MIN_WITHIN_TOL, float(d.max())
    stats.update(agree=agree, max_px=float(d.max()))
    if boxes is not None:
        ious = []
        for t in range(len(boxes)):
            s = slice(off[t], off[t + 1])
            b, _ = oracle.local_search_klt(boxes[t], pts[s], c_o[s], c_st[s])
            ious.append(cv2_ref.box_iou(nb[t], b))
        assert min(ious) >= MIN_IOU, min(ious)
        stats["min_iou"] = float(min(ious))
    return stats


def test_c1_rig_vs_oracle_and_cv2(oracle):
    w, h, ntgt, per, levels = 768, 576, 20, 100, 3
    cam = synth.Camera(w, h, ntgt, cam=11)
    f0, f1 = cam.frame(0), cam.frame(1)
    boxes = cam.boxes(0)
    with klt.KLTContext(w, h, levels=levels, fb_thresh=1.0) as ctx:
        ctx.push_frame(f0)
        seeds, counts = ctx.detect(boxes, per, 0.01, 3.0)          # Shi-Tomasi seeds inside the boxes, as bench.py's step
        ctx.push_frame(f1)
        pts = seeds.reshape(-1, 2)
        off = (np.arange(ntgt + 1) * per).astype(np.int32)
        stats = _compare(ctx, oracle, f0, f1, pts, levels, boxes, off)
    assert stats["tracked"] > 0.5


def test_c3_rig_vs_oracle_and_cv2(oracle):
    """1920x1080, 200 targets x 256 seeds = 51 200 seeds, 4 levels."""
    w, h, ntgt, per, levels = 1920, 1080, 200, 256, 4
    cam = synth.Camera(w, h, ntgt, cam=12)
    f0, f1 = cam.frame(0), cam.frame(1)
    boxes = cam.boxes(0)
    pts, tid = synth.seeds_in_boxes(boxes, per, w, h, np.random.default_rng(77))
    off = (np.arange(ntgt + 1) * per).astype(np.int32)
    with klt.KLTContext(w, h, levels=levels, fb_thresh=1.0, max_points=1 << 16) as ctx:
        ctx.push_frame(f0); ctx.push_frame(f1)
        stats = _compare(ctx, oracle, f0, f1, pts, levels, boxes, off)
    assert stats["tracked"] > 0.5


def test_c3_detected_seeds_equal_oracle(oracle):
    """C3's detection leg: 256 Shi-Tomasi seeds in each of 200 boxes of a 1080p frame, identical lists."""
    w, h, ntgt, per = 1920, 1080, 200, 256
    cam = synth.Camera(w, h, ntgt, cam=13)
    f0 = cam.frame(0)
    boxes = cam.boxes(0)
    with klt.KLTContext(w, h, levels=4, max_targets=256) as ctx:
        ctx.push_frame(f0)
        seeds, counts = ctx.detect(boxes, per, 0.01, 3.0)
        eig = ctx.read_min_eig()
    assert np.array_equal(eig, oracle.min_eig_map(f0))
    for t in range(ntgt):
        want = oracle.detect(f0, boxes[t], per, 0.01, 3.0, eig=eig)
        assert counts[t] == len(want) and np.array_equal(seeds[t, :counts[t]], want), t


def test_c4_subsample_vs_oracle_and_cv2(oracle):
    """3840x2160, 5 levels, every 10th seed of the stride-4 grid (51 840 seeds) + the grid's border rows and columns."""
    w, h, levels = 3840, 2160, 5
    cam = synth.Camera(w, h, 0, cam=14)
    f0, f1 = cam.frame(0), cam.frame(1)
    grid = synth.grid_seeds(w, h, 4, 2)
    edge = (grid[:, 0] < 6) | (grid[:, 0] > w - 7) | (grid[:, 1] < 6) | (grid[:, 1] > h - 7)
    pts = np.ascontiguousarray(np.concatenate([grid[::10], grid[edge][::3]]))
    assert len(pts) >= 50000
    with klt.KLTContext(w, h, levels=levels, fb_thresh=1.0, max_points=len(pts) + 16) as ctx:
        ctx.push_frame(f0); ctx.push_frame(f1)
        assert ctx.levels == 5
        stats = _compare(ctx, oracle, f0, f1, pts, levels)
    assert stats["tracked"] > 0.9


def test_untrackable_window_disables_only_its_target(oracle):
    """A degenerate / oversize per-target window (the reference passes (int)box.w x (int)box.h unchecked, Tracker2D.cpp:782,
    :877) must not void the batch: that target's points come back unmoved with status 0, the others equal the oracle."""
    w, h, ntgt = 320, 240, 5
    cam = synth.Camera(w, h, ntgt, cam=15)
    f0, f1 = cam.frame(0), cam.frame(1)
    pts, tid = synth.seeds_in_boxes(cam.boxes(0), 40, w, h, np.random.default_rng(2))
    win = np.array([[21, 21], [0, 30], [25, 60], [400, 30], [31, 240]], np.int32)     # targets 1, 3, 4 cannot be tracked
    pa, pb = oracle.build_pyramid(f0, 3), oracle.build_pyramid(f1, 3)
    for fb_thresh in (-1.0, 1.0):
        with klt.KLTContext(w, h, levels=4, fb_thresh=fb_thresh) as ctx:
            ctx.push_frame(f0); ctx.push_frame(f1)
            o, st, er, fb = ctx.track(pts, target_of_pt=tid, win_wh=win)
        for t in range(ntgt):
            m = tid == t
            if t in (1, 3, 4):
                assert st[m].sum() == 0 and np.array_equal(o[m], pts[m]) and not er[m].any()
            elif fb_thresh < 0:
                g_o, g_st, g_er = oracle.lk_track(pa, pb, pts[m], (int(win[t, 0]), int(win[t, 1])), 3, exact_sums=True)
                assert np.array_equal(st[m], g_st) and np.abs(o[m] - g_o).max() <= 1e-4
            else:
                g_o, g_st, g_er, g_fb = oracle.track_fb(pa, pb, pts[m], (int(win[t, 0]), int(win[t, 1])), 3, fb_thresh=1.0, exact_sums=True)
                assert np.array_equal(st[m], g_st) and np.abs(o[m] - g_o).max() <= 1e-4
    # all-21x21 batch (persistent kernel) on a frame narrower than the window
    small = np.ascontiguousarray(f0[:64, :20])
    with klt.KLTContext(20, 64, levels=2, fb_thresh=1.0) as ctx:
        ctx.push_frame(small); ctx.push_frame(small)
        p = np.float32([[5, 5], [10, 30]])
        o, st, er, fb = ctx.track(p)
        assert st.sum() == 0 and np.array_equal(o, p)


def _pinned(shape, dtype):
    torch = pytest.importorskip("torch")
    t = torch.zeros(shape, dtype=getattr(torch, np.dtype(dtype).name)).pin_memory()
    return t, t.numpy()


def test_pinned_memory_path_equals_host_path(oracle):
    """PSNW_MEM_PINNED: every input and output of detect / track_batch / box_update / points_in_boxes in page-locked host
    memory, calls return at once, results valid after psnw_sync -- identical to the synchronous PSNW_MEM_HOST results."""
    import ctypes as C
    w, h, ntgt, per = 768, 576, 24, 100
    cam = synth.Camera(w, h, ntgt, cam=16)
    f0, f1 = cam.frame(0), cam.frame(1)
    boxes = cam.boxes(0)
    n = ntgt * per
    off = (np.arange(ntgt + 1) * per).astype(np.int32)
    with klt.KLTContext(w, h, levels=4, fb_thresh=1.0) as ctx:
        keep = []
        def pin(shape, dt):
            t, a = _pinned(shape, dt); keep.append(t); return a
        pf0, pf1 = pin((h, w), np.uint8), pin((h, w), np.uint8)
        pf0[:] = f0; pf1[:] = f1
        ctx.push_frames([pf0.ctypes.data], mem=klt.MEM_PINNED, pitch=w)
        p_seeds, p_cnt = pin((ntgt, per, 2), np.float32), pin((ntgt,), np.int32)
        ctx._ck(ctx.lib.psnw_klt_detect(ctx._h, 0, 0, boxes.ctypes.data, ntgt, per, 0.01, 3.0, p_seeds.ctypes.data, p_cnt.ctypes.data, klt.MEM_PINNED))
        ctx.sync()
        h_seeds, h_cnt = ctx.detect(boxes, per, 0.01, 3.0)
        assert np.array_equal(p_seeds, h_seeds) and np.array_equal(p_cnt, h_cnt)
        ctx.push_frames([pf1.ctypes.data], mem=klt.MEM_PINNED, pitch=w)
        p_out, p_st, p_er, p_fb = pin((n, 2), np.float32), pin((n,), np.uint8), pin((n,), np.float32), pin((n,), np.float32)
        pts = p_seeds.reshape(-1, 2)
        ctx.track_batch([dict(cam=0, slot_from=1, slot_to=0, n_pts=n, pts=pts, out=p_out, status=p_st, err=p_er, fb_err=p_fb)], mem=klt.MEM_PINNED)
        p_box, p_mask, p_ic = pin((ntgt, 4), np.float64), pin((n,), np.uint8), pin((ntgt,), np.int32)
        p_boxes_in, p_off = pin((ntgt, 4), np.float64), pin((ntgt + 1,), np.int32)
        p_boxes_in[:] = boxes; p_off[:] = off
        ctx.sync()                        # (box_update reads p_out / p_st on the host side of the copy: they must have landed)
        ctx.box_update_raw(p_boxes_in.ctypes.data, ntgt, pts.ctypes.data, p_out.ctypes.data, p_st.ctypes.data, p_off.ctypes.data,
                           p_box.ctypes.data, p_mask.ctypes.data, p_ic.ctypes.data, klt.MEM_PINNED)
        p_pib = pin((ntgt, ntgt), np.int32)
        ctx._ck(ctx.lib.psnw_points_in_boxes(ctx._h, p_out.ctypes.data, p_off.ctypes.data, ntgt, p_boxes_in.ctypes.data, ntgt,
                                             p_pib.ctypes.data, klt.MEM_PINNED))
        ctx.sync()
        o, st, er, fb = ctx.track(pts)
        assert np.array_equal(p_out, o) and np.array_equal(p_st, st) and np.array_equal(p_er, er) and np.array_equal(p_fb, fb)
        nb, mask, cnt = ctx.box_update(boxes, pts, o, off, status=st)
        assert np.array_equal(p_box, nb) and np.array_equal(p_mask, mask) and np.array_equal(p_ic, cnt)
        assert np.array_equal(p_pib, ctx.points_in_boxes(o, off, boxes))
    # and the host results themselves equal the oracle
    pa, pb = oracle.build_pyramid(f0, 3), oracle.build_pyramid(f1, 3)
    g_o, g_st, g_er, g_fb = oracle.track_fb(pa, pb, pts, fb_thresh=1.0, exact_sums=True)
    assert np.array_equal(st, g_st) and np.abs(o - g_o).max() <= 1e-4


def test_failed_push_leaves_the_ring_untouched():
    """A push that fails after the ring has been rotated (null frame among the BGR pointers, bad camera) must not leave slot 0
    pointing at a recycled plane."""
    import ctypes as C
    w, h = 96, 64
    rng = np.random.default_rng(0)
    a, b = rng.integers(0, 256, (h, w), np.uint8), rng.integers(0, 256, (h, w), np.uint8)
    with klt.KLTContext(w, h, n_cameras=2, levels=2) as ctx:
        ctx.push_frames([a, b])
        before = [ctx.read_level(c, 0, 0)[0].copy() for c in range(2)]
        bgr = np.ascontiguousarray(np.repeat(a[:, :, None], 3, axis=2))
        ptrs = (C.c_void_p * 2)(bgr.ctypes.data, None)
        cams = np.arange(2, dtype=np.int32)
        rc = ctx.lib.psnw_push_frames_bgr(ctx._h, 2, cams.ctypes.data, ptrs, w * 3, 0, 15, klt.MEM_HOST)
        assert rc == klt.EARG
        bad = np.array([0, 5], np.int32)
        ptrs2 = (C.c_void_p * 2)(a.ctypes.data, b.ctypes.data)
        assert ctx.lib.psnw_push_frames(ctx._h, 2, bad.ctypes.data, ptrs2, w, 0, klt.MEM_HOST) == klt.EARG
        assert ctx.frames_held(0) == 1 and ctx.frames_held(1) == 1
        for c in range(2):
            assert np.array_equal(ctx.read_level(c, 0, 0)[0], before[c])
