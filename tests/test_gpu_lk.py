"""GPU tier: K2/K3 (pyramidal LK + fused forward-backward check) through the C ABI.

Gates (north_star): status agreement >= 99.5 %, position within 0.05 px for >= 99 % of agreeing points -- against
the cv2 fixtures; and a much tighter gate against the oracle in exact-sum mode, which performs the same
arithmetic as the kernel (identical status, positions within 1e-4 px)."""
import numpy as np
import pytest

from mcmtt_opticalflow_b200 import klt, synth

pytestmark = pytest.mark.gpu

TOL_PX = 0.05          # north_star position tolerance
MIN_STATUS_AGREE = 0.995
MIN_WITHIN_TOL = 0.99


def _gate(o, st, g_o, g_st, tag=""):
    assert (st == g_st).mean() >= MIN_STATUS_AGREE, tag
    ok = (st == 1) & (g_st == 1)
    if ok.any():
        d = np.linalg.norm(o - g_o, axis=1)[ok]
        assert (d <= TOL_PX).mean() >= MIN_WITHIN_TOL, (tag, float(d.max()))


def _tight(o, st, er, g_o, g_st, g_er, tag=""):
    assert np.array_equal(st, g_st), (tag, int((st != g_st).sum()))
    d = np.linalg.norm(o - g_o, axis=1)
    assert d.max() <= 1e-4, (tag, float(d.max()))
    assert np.abs(er - g_er).max() <= 1e-5, tag


def test_golden_single_direction(oracle, golden):
    a, b, pts = golden["a"], golden["b"], golden["lk_pts"]
    for i, (ww, wh, ml) in enumerate(golden["lk_cfgs"]):
        ww, wh, ml = int(ww), int(wh), int(ml)
        with klt.KLTContext(a.shape[1], a.shape[0], levels=ml + 1, win=(ww, wh), fb_thresh=-1.0) as ctx:
            ctx.push_frame(a); ctx.push_frame(b)
            o, st, er, _ = ctx.track(pts, slot_from=1, slot_to=0)
        _gate(o, st, golden["lk_out_%d" % i], golden["lk_st_%d" % i], (ww, wh, ml))
        pa, pb = oracle.build_pyramid(a, ml), oracle.build_pyramid(b, ml)
        g_o, g_st, g_er = oracle.lk_track(pa, pb, pts, (ww, wh), ml, exact_sums=True)
        _tight(o, st, er, g_o, g_st, g_er, (ww, wh, ml))


def test_golden_forward_backward(oracle, golden):
    a, b, pts = golden["a"], golden["b"], golden["lk_pts"]
    with klt.KLTContext(a.shape[1], a.shape[0], levels=4, fb_thresh=1.0) as ctx:
        ctx.push_frame(a); ctx.push_frame(b)
        o, st, er, fb = ctx.track(pts)
    _gate(o, st, golden["fb_out"], golden["fb_st"], "fb")
    ok = (st == 1) & (golden["fb_st"] == 1)
    assert np.abs(fb - golden["fb_err"])[ok].max() < TOL_PX
    pa, pb = oracle.build_pyramid(a, 3), oracle.build_pyramid(b, 3)
    g_o, g_st, g_er, g_fb = oracle.track_fb(pa, pb, pts, fb_thresh=1.0, exact_sums=True)
    _tight(o, st, er, g_o, g_st, g_er, "fb")
    assert np.abs(fb - g_fb).max() <= 1e-4


def test_odd_frame_size(oracle, golden):
    a, b, pts = golden["pyr_odd_0"], golden["odd_b"], golden["lk_pts"]
    with klt.KLTContext(a.shape[1], a.shape[0], levels=4, fb_thresh=-1.0) as ctx:
        ctx.push_frame(a); ctx.push_frame(b)
        o, st, er, _ = ctx.track(pts)
    _gate(o, st, golden["lk_odd_out"], golden["lk_odd_st"], "odd")


def test_generic_and_w21_kernels_agree(golden):
    """The 21x21 fast path and the generic kernel are the same arithmetic: results must be bit-identical."""
    a, b, pts = golden["a"], golden["b"], golden["lk_pts"]
    n = len(pts)
    with klt.KLTContext(a.shape[1], a.shape[0], levels=4, fb_thresh=1.0) as ctx:
        ctx.push_frame(a); ctx.push_frame(b)
        o1, st1, er1, fb1 = ctx.track(pts)                                        # all 21x21 -> lk_w21_kernel
        tgt = np.zeros(n, np.int32); tgt[-1] = 1
        o2, st2, er2, fb2 = ctx.track(pts, target_of_pt=tgt, win_wh=[[21, 21], [23, 23]])   # mixed -> lk_generic_kernel
    assert np.array_equal(o1[:-1], o2[:-1]) and np.array_equal(st1[:-1], st2[:-1])
    assert np.array_equal(er1[:-1], er2[:-1]) and np.array_equal(fb1[:-1], fb2[:-1])


@pytest.mark.parametrize("fb", [False, True])
def test_c2_shaped_batch_vs_oracle(oracle, fb):
    """BASELINE config C2 shape: 4 views 768x576, 40 targets x 100 seeds per view, one batched launch."""
    w, h, ncam, ntgt, per = 768, 576, 4, 40, 100
    cams = [synth.Camera(w, h, ntgt, cam=c, low_texture=(c == 3)) for c in range(ncam)]
    with klt.KLTContext(w, h, n_cameras=ncam, levels=4, fb_thresh=1.0 if fb else -1.0, max_points=1 << 15) as ctx:
        f0 = [c.frame(0) for c in cams]; f1 = [c.frame(1) for c in cams]
        ctx.push_frames(f0); ctx.push_frames(f1)
        jobs, keep = [], []
        for c in range(ncam):
            pts, tid = synth.seeds_in_boxes(cams[c].boxes(0), per, w, h, np.random.default_rng(100 + c))
            n = len(pts)
            j = dict(cam=c, slot_from=1, slot_to=0, n_pts=n, pts=pts, out=np.zeros((n, 2), np.float32),
                     status=np.zeros(n, np.uint8), err=np.zeros(n, np.float32), fb_err=np.zeros(n, np.float32))
            jobs.append(j); keep.append(pts)
        ctx.track_batch(jobs)
    agree, within, total = 0, 0, 0
    for c in range(ncam):
        pa, pb = oracle.build_pyramid(f0[c], 3), oracle.build_pyramid(f1[c], 3)
        if fb:
            g_o, g_st, g_er, g_fb = oracle.track_fb(pa, pb, keep[c], exact_sums=True)
        else:
            g_o, g_st, g_er = oracle.lk_track(pa, pb, keep[c], exact_sums=True)
        j = jobs[c]
        _tight(j["out"], j["status"], j["err"], g_o, g_st, g_er, ("cam", c))
        if fb:
            assert np.abs(j["fb_err"] - g_fb).max() <= 1e-4
        assert j["status"].mean() > 0.5          # the synthetic texture is trackable


@pytest.mark.parametrize("apron", [0, 64])
@pytest.mark.parametrize("w,h", [(200, 150), (333, 97), (768, 576)])
def test_border_seeds_all_levels(oracle, w, h, apron):
    """Seeds hugging every border and corner: the windows cross the frame edge on every pyramid level (REFLECT_101 image
    apron kept in HBM on levels >= 1, on-the-fly mirror on level 0, zero derivative apron), some leave the frame."""
    rng = np.random.default_rng(w + h)
    cam = synth.Camera(w, h, 3, cam=4)
    f0, f1 = cam.frame(0), cam.frame(1)
    n = 600
    t = rng.uniform(0, 1, n)
    band = rng.uniform(-3.0, 14.0, n)                       # distance from the edge, a few seeds outside the frame
    side = rng.integers(0, 4, n)
    x = np.where(side == 0, band, np.where(side == 1, w - 1 - band, t * (w - 1)))
    y = np.where(side == 2, band, np.where(side == 3, h - 1 - band, t * (h - 1)))
    corners = np.array([[0.5, 0.5], [w - 1.5, 0.5], [0.5, h - 1.5], [w - 1.5, h - 1.5], [10.0, 10.0], [w - 11.0, h - 11.0]])
    pts = np.concatenate([np.stack([x, y], 1), corners]).astype(np.float32)
    with klt.KLTContext(w, h, levels=4, fb_thresh=1.0, apron=apron) as ctx:
        ctx.push_frame(f0); ctx.push_frame(f1)
        o, st, er, fb = ctx.track(pts)
        ml = ctx.levels - 1
    pa, pb = oracle.build_pyramid(f0, ml), oracle.build_pyramid(f1, ml)
    g_o, g_st, g_er, g_fb = oracle.track_fb(pa, pb, pts, (21, 21), ml, fb_thresh=1.0, exact_sums=True)
    _tight(o, st, er, g_o, g_st, g_er, ("border", w, h))
    assert np.abs(fb - g_fb).max() <= 1e-4
    assert st.mean() > 0.05


def test_ragged_many_jobs_batch(oracle):
    """One launch over 11 jobs -- more than fit in the kernel parameters, so the descriptors travel through the staging
    path -- on two cameras and several slot pairs, with empty jobs in between and ragged sizes."""
    w, h = 320, 240
    cams = [synth.Camera(w, h, 5, cam=c) for c in range(2)]
    rng = np.random.default_rng(3)
    with klt.KLTContext(w, h, n_cameras=2, levels=4, fb_thresh=1.0, max_points=1 << 13) as ctx:
        frames = [[c.frame(k) for k in range(3)] for c in cams]
        for k in range(3):
            ctx.push_frames([frames[0][k], frames[1][k]])
        sizes = [0, 1, 37, 0, 256, 5, 64, 0, 100, 33, 2]
        jobs = []
        for j, n in enumerate(sizes):
            cam, (sf, st_) = j % 2, [(1, 0), (2, 1), (2, 0), (0, 1)][j % 4]
            pts = rng.uniform(-5, [w + 5, h + 5], size=(n, 2)).astype(np.float32)
            jobs.append(dict(cam=cam, slot_from=sf, slot_to=st_, n_pts=n, pts=pts, out=np.zeros((n, 2), np.float32),
                             status=np.zeros(n, np.uint8), err=np.zeros(n, np.float32), fb_err=np.zeros(n, np.float32)))
        ctx.track_batch(jobs)
    pyr = [[oracle.build_pyramid(f, 3) for f in fr] for fr in frames]
    for j in jobs:
        if j["n_pts"] == 0:
            continue
        a, b = pyr[j["cam"]][2 - j["slot_from"]], pyr[j["cam"]][2 - j["slot_to"]]      # slot k = k frames ago
        g_o, g_st, g_er, g_fb = oracle.track_fb(a, b, j["pts"], fb_thresh=1.0, exact_sums=True)
        _tight(j["out"], j["status"], j["err"], g_o, g_st, g_er, ("job", j["n_pts"]))
        assert np.abs(j["fb_err"] - g_fb).max() <= 1e-4


@pytest.mark.parametrize("apron", [0, 64])
def test_reference_faithful_per_target_windows(oracle, apron):
    """Reference call pattern: window (int)box.w x (int)box.h per target (Tracker2D.cpp:877), default 4 levels."""
    w, h, ntgt = 768, 576, 12
    cam = synth.Camera(w, h, ntgt, cam=9)
    f0, f1 = cam.frame(0), cam.frame(1)
    boxes = cam.boxes(0)
    pts, tid = synth.seeds_in_boxes(boxes, 50, w, h, np.random.default_rng(5))
    win = np.stack([boxes[:, 2].astype(np.int32), boxes[:, 3].astype(np.int32)], 1)
    with klt.KLTContext(w, h, levels=4, fb_thresh=-1.0, apron=apron) as ctx:
        ctx.push_frame(f0); ctx.push_frame(f1)
        o, st, er, _ = ctx.track(pts, target_of_pt=tid, win_wh=win)
    pa, pb = oracle.build_pyramid(f0, 3), oracle.build_pyramid(f1, 3)
    for t in range(ntgt):
        m = tid == t
        g_o, g_st, g_er = oracle.lk_track(pa, pb, pts[m], (int(win[t, 0]), int(win[t, 1])), 3, exact_sums=True)
        _tight(o[m], st[m], er[m], g_o, g_st, g_er, ("target", t, win[t]))


def test_edge_cases(oracle):
    w, h = 160, 120
    rng = np.random.default_rng(0)
    a = synth.Camera(w, h, 0, cam=1).frame(0)
    flat = np.full((h, w), 128, np.uint8)
    with klt.KLTContext(w, h, levels=3, fb_thresh=-1.0) as ctx:
        ctx.push_frame(a); ctx.push_frame(a)
        o, st, er, _ = ctx.track(np.zeros((0, 2), np.float32))                   # empty input
        assert o.shape == (0, 2)
        far = np.float32([[-500, -500], [1e6, 5], [w + 40, h + 40], [-11, 5], [w + 9.5, h / 2]])
        o, st, er, _ = ctx.track(far)
        pa = oracle.build_pyramid(a, 2)
        g_o, g_st, g_er = oracle.lk_track(pa, pa, far, (21, 21), 2, exact_sums=True)
        assert np.array_equal(st, g_st) and np.allclose(o, g_o, atol=1e-4)
        pts = rng.uniform(0, [w, h], size=(64, 2)).astype(np.float32)            # identical frames: zero motion
        o, st, er, _ = ctx.track(pts)
        assert np.abs(o - pts)[st == 1].max() < 1e-3
        ctx.push_frame(flat); ctx.push_frame(flat)                               # no texture: min-eig gate rejects all
        o, st, er, _ = ctx.track(pts)
        assert st.sum() == 0
        o, st, er, _ = ctx.track(pts, win_wh=[[w, 5]])                             # window >= frame: target skipped, not an error
        assert st.sum() == 0 and np.array_equal(o, pts) and not er.any()
    with klt.KLTContext(w, h, levels=3) as ctx:
        with pytest.raises(klt.PsnwError) as e:
            ctx.track(np.zeros((4, 2), np.float32))                              # ring empty
        assert e.value.code == klt.ESTATE


def test_linearity_property_at_full_size():
    """Size-independent property at C3 size (1080p, 51 200 seeds): a frame translated by an integer offset must be
    tracked to p + offset away from the borders, and tracking a frame onto itself must return p."""
    w, h = 1920, 1080
    cam = synth.Camera(w, h, 0, cam=2)
    a = cam.frame(0)
    dx, dy = 3, -2
    b = np.roll(np.roll(a, dy, axis=0), dx, axis=1)
    rng = np.random.default_rng(1)
    pts = rng.uniform([60, 60], [w - 60, h - 60], size=(51200, 2)).astype(np.float32)
    with klt.KLTContext(w, h, levels=4, fb_thresh=1.0, max_points=1 << 16) as ctx:
        ctx.push_frame(a); ctx.push_frame(b)
        o, st, er, fb = ctx.track(pts)
        assert st.mean() > 0.97
        d = o[st == 1] - pts[st == 1] - np.float32([dx, dy])
        assert np.percentile(np.abs(d), 99) < 0.05
        assert (fb[st == 1] <= 1.0).all()
        ctx.push_frame(b)
        o2, st2, _, _ = ctx.track(pts)
        assert np.abs(o2 - pts)[st2 == 1].max() < 1e-3


def test_c4_dense_grid_properties_at_4k():
    """BASELINE config C4 at full size (4K, stride-4 grid = 518 400 seeds, 5 levels): size-independent properties --
    a frame tracked onto itself returns the seeds (idempotence), an integer-translated frame returns seeds + offset,
    the FB error of validated points respects the gate, and the batched result does not depend on how the seeds are
    split into jobs (sharding invariance)."""
    w, h = 3840, 2160
    cam = synth.Camera(w, h, 0, cam=4)
    a = cam.frame(0)
    dx, dy = -2, 3
    b = np.roll(np.roll(a, dy, axis=0), dx, axis=1)
    pts = synth.grid_seeds(w, h, 4, 2)
    n = len(pts)
    assert n == 518400
    with klt.KLTContext(w, h, levels=5, fb_thresh=1.0, max_points=n + 16) as ctx:
        ctx.push_frame(a); ctx.push_frame(a)
        o, st, er, fb = ctx.track(pts)
        assert st.mean() > 0.99 and np.abs(o - pts)[st == 1].max() < 1e-3 and er[st == 1].max() < 1e-3
        ctx.push_frame(b)
        o, st, er, fb = ctx.track(pts)
        inner = (pts[:, 0] > 40) & (pts[:, 0] < w - 40) & (pts[:, 1] > 40) & (pts[:, 1] < h - 40)
        ok = (st == 1) & inner
        assert ok.sum() > 0.95 * inner.sum()
        d = o[ok] - pts[ok] - np.float32([dx, dy])
        assert np.percentile(np.abs(d), 99.5) < 0.05
        assert (fb[st == 1] <= 1.0).all()
        # the same seeds as three jobs of unequal size in one batched launch: bit-identical outputs
        cuts = [0, 100001, 350000, n]
        jobs = []
        for i in range(3):
            m = cuts[i + 1] - cuts[i]
            jobs.append(dict(cam=0, slot_from=1, slot_to=0, n_pts=m, pts=np.ascontiguousarray(pts[cuts[i]:cuts[i + 1]]),
                             out=np.zeros((m, 2), np.float32), status=np.zeros(m, np.uint8), err=np.zeros(m, np.float32),
                             fb_err=np.zeros(m, np.float32)))
        ctx.track_batch(jobs)
        assert np.array_equal(np.concatenate([j["out"] for j in jobs]), o)
        assert np.array_equal(np.concatenate([j["status"] for j in jobs]), st)


def test_north_star_gates_against_cv2_live(oracle):
    """north_star acceptance measured directly against the library the reference links, when cv2 is importable on the box
    (same image as the build container): C2-shaped rig, identical frames and seeds, forward + backward LK + FB gate and
    the LocalSearchKLT boxes.  Status agreement >= 99.5 %, <= 0.05 px for >= 99 % of agreeing points, box IoU >= 0.99."""
    cv2_ref = pytest.importorskip("oracle.cv2_ref")
    w, h, ncam, ntgt, per = 768, 576, 4, 40, 100
    cams = [synth.Camera(w, h, ntgt, cam=c, low_texture=(c == 3)) for c in range(ncam)]
    f0 = [c.frame(0) for c in cams]; f1 = [c.frame(1) for c in cams]
    agree = within = n_ok = total = 0
    ious = []
    with klt.KLTContext(w, h, n_cameras=ncam, levels=4, fb_thresh=1.0, max_points=1 << 15) as ctx:
        ctx.push_frames(f0); ctx.push_frames(f1)
        for c in range(ncam):
            boxes = cams[c].boxes(0)
            pts, tid = synth.seeds_in_boxes(boxes, per, w, h, np.random.default_rng(200 + c))
            o, st, er, fb = ctx.track(pts, cam=c)
            g_o, g_st, g_er, g_fb = cv2_ref.track_fb(f0[c], f1[c], pts, fb_thresh=1.0)
            agree += int((st == g_st).sum()); total += len(pts)
            ok = (st == 1) & (g_st == 1)
            d = np.linalg.norm(o - g_o, axis=1)[ok]
            within += int((d <= TOL_PX).sum()); n_ok += int(ok.sum())
            off = np.concatenate([[0], np.cumsum(np.bincount(tid, minlength=ntgt))]).astype(np.int32)
            nb, mask, cnt = ctx.box_update(boxes, pts, o, off, status=st)
            for t in range(ntgt):
                s = slice(off[t], off[t + 1])
                b, _ = oracle.local_search_klt(boxes[t], pts[s], g_o[s], g_st[s])
                ious.append(cv2_ref.box_iou(nb[t], b))
    assert agree / total >= MIN_STATUS_AGREE, agree / total
    assert within / max(n_ok, 1) >= MIN_WITHIN_TOL, within / max(n_ok, 1)
    assert np.min(ious) >= 0.99, np.min(ious)                # measured: 1.0 / 1.0 (max 0.006 px) / min IoU 0.999998


def test_split_tail_items_equal_fused_items(monkeypatch):
    """The persistent 21x21 kernel tracks the last points of a large batch as two work items (forward pass, then backward
    pass, handed over through a flag) so that the grid drains sooner.  Whatever the length of that tail -- none, a few
    points, half the batch -- and however often the launch repeats (the flags clean themselves), all outputs are
    bit-identical."""
    w, h = 768, 576
    cam = synth.Camera(w, h, 30, cam=3)
    a, b = cam.frame(0), cam.frame(1)
    rng = np.random.default_rng(23)
    pts = np.concatenate([synth.seeds_in_boxes(cam.boxes(0), 200, w, h, rng)[0],
                          rng.uniform(-5, [w + 5, h + 5], size=(300, 2)).astype(np.float32)])     # some seeds leave the frame
    res = {}
    for tail in ("0", "7", "1000", str(len(pts) // 2)):
        monkeypatch.setenv("PSNW_LK_TAIL", tail)
        with klt.KLTContext(w, h, levels=4, fb_thresh=1.0, max_points=len(pts)) as ctx:
            ctx.push_frame(a); ctx.push_frame(b)
            runs = [ctx.track(pts) for _ in range(3)]
        for r in runs[1:]:
            for x, y in zip(runs[0], r):
                assert np.array_equal(x, y), tail
        res[tail] = runs[0]
    for tail, r in res.items():
        for x, y in zip(res["0"], r):
            assert np.array_equal(x, y), tail
    assert 0.5 < res["0"][1].mean() < 1.0
