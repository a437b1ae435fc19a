"""GPU tier: K4 (LocalSearchKLT box motion, bit-exact in double) and K5 (Shi-Tomasi detect) through the C ABI."""
import numpy as np
import pytest

from mcmtt_opticalflow_b200 import klt, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = klt.KLTContext(256, 192, levels=3, max_points=1 << 17, max_targets=1 << 10)
    yield c
    c.close()


def test_box_update_golden(ctx, golden):
    off = np.concatenate([[0], np.cumsum(golden["ls_n"])]).astype(np.int32)
    ioff = np.concatenate([[0], np.cumsum(golden["ls_inl_n"])])
    boxes, mask, counts = ctx.box_update(golden["ls_box"], golden["ls_pre"], golden["ls_cur"], off)
    assert np.array_equal(boxes, golden["ls_box_out"])           # bit-exact doubles
    assert np.array_equal(counts, golden["ls_inl_n"])
    for c in range(len(golden["ls_n"])):
        inl = np.flatnonzero(mask[off[c]:off[c + 1]])
        assert np.array_equal(inl, golden["ls_inl"][ioff[c]:ioff[c + 1]]), c


def test_box_update_random_vs_oracle(ctx, oracle):
    rng = np.random.default_rng(42)
    n_t = 200
    ns = rng.integers(0, 257, size=n_t)
    ns[:3] = [0, 1, 256]
    off = np.concatenate([[0], np.cumsum(ns)]).astype(np.int32)
    pre = rng.uniform(0, 1900, size=(off[-1], 2)).astype(np.float32)
    cur = pre.copy()
    boxes = np.stack([rng.uniform(0, 1800, n_t), rng.uniform(0, 900, n_t), rng.uniform(5, 140, n_t), rng.uniform(20, 330, n_t)], 1)
    for t in range(n_t):
        s = slice(off[t], off[t + 1])
        kind = t % 4
        d = rng.uniform(-8, 8, size=2)
        if kind == 0:
            cur[s] = pre[s] + d + rng.normal(0, 0.4, size=(ns[t], 2))
        elif kind == 1:
            cur[s] = pre[s] + np.round(rng.uniform(-2, 2, size=(ns[t], 2)) * 4) / 4       # many exact ties
        elif kind == 2:
            cur[s] = pre[s] + rng.normal(0, 0.06, size=(ns[t], 2))                        # around the 0.1 px gate
        else:
            cur[s] = pre[s] + d
            k = ns[t] // 2
            cur[off[t]:off[t] + k] = pre[off[t]:off[t] + k] + rng.uniform(-40, 40, size=(k, 2))
    cur = cur.astype(np.float32)
    status = (rng.random(off[-1]) < 0.8).astype(np.uint8)
    for st in (None, status):
        g_boxes, g_mask, g_counts = ctx.box_update(boxes, pre, cur, off, status=st)
        for t in range(n_t):
            s = slice(off[t], off[t + 1])
            b, inl = oracle.local_search_klt(boxes[t], pre[s], cur[s], None if st is None else st[s])
            assert np.array_equal(g_boxes[t], b), t
            assert np.array_equal(np.flatnonzero(g_mask[s]), inl), t
            assert g_counts[t] == len(inl)


def test_box_update_maximum_target_size(ctx, oracle):
    """PSNW_MAX_PTS_PER_TARGET = 4096 features in one target (the capacity of the shared-memory arrays), next to a
    3000-feature and an empty one."""
    rng = np.random.default_rng(7)
    ns = np.array([4096, 0, 3000])
    off = np.concatenate([[0], np.cumsum(ns)]).astype(np.int32)
    pre = rng.uniform(0, 1900, size=(off[-1], 2)).astype(np.float32)
    cur = (pre + rng.uniform(-3, 3, size=2) + np.round(rng.normal(0, 1.0, size=pre.shape) * 8) / 8).astype(np.float32)
    boxes = np.array([[100.0, 80.0, 61.5, 140.0], [5.0, 5.0, 40.0, 96.0], [700.0, 300.0, 33.25, 80.0]])
    status = (rng.random(off[-1]) < 0.9).astype(np.uint8)
    for st in (None, status):
        g_boxes, g_mask, g_counts = ctx.box_update(boxes, pre, cur, off, status=st)
        for t in range(3):
            s = slice(off[t], off[t + 1])
            b, inl = oracle.local_search_klt(boxes[t], pre[s], cur[s], None if st is None else st[s])
            assert np.array_equal(g_boxes[t], b), t
            assert np.array_equal(np.flatnonzero(g_mask[s]), inl), t
            assert g_counts[t] == len(inl)


def test_min_eig_map_bit_exact_and_selection(golden, oracle):
    a = golden["a"]
    with klt.KLTContext(a.shape[1], a.shape[0], levels=2) as c:
        c.push_frame(a)
        eig = c.read_min_eig()
        assert np.array_equal(eig, oracle.min_eig_map(a))                         # same float ops, same order
        assert np.abs(eig - golden["eig_a"]).max() <= 2e-6 * golden["eig_a"].max()
        boxes = golden["gftt_boxes"]
        for si, (mc, q, md) in enumerate(golden["gftt_cfgs"]):
            pts, counts = c.detect(boxes, int(mc), float(q), float(md))
            for bi in range(len(boxes)):
                got = pts[bi, :counts[bi]]
                want = oracle.detect(a, boxes[bi], int(mc), float(q), float(md), eig=eig)
                assert got.shape == want.shape and np.array_equal(got, want), (bi, si)
                g = golden["gftt_%d_%d" % (bi, si)]                               # cv2: own map -> near-ties may reorder
                assert len(set(map(tuple, got)) ^ set(map(tuple, g))) <= max(2, len(g) // 50), (bi, si)
    t = golden["tie_img"]
    with klt.KLTContext(64, 64, levels=1) as c:
        c.push_frame(t)
        pts, counts = c.detect([[0, 0, 200, 200]], 50, 0.01, 3.0)
        assert np.array_equal(pts[0, :counts[0]], golden["tie_out"])               # exact ties: higher address first


@pytest.mark.parametrize("box,max_per,min_dist", [([0, 0, 768, 576], 300, 4.0),      # candidates exceed the shared list
                                                  ([100, 80, 400, 300], 1500, 2.0),   # more corners than the shared accepted list
                                                  ([10, 10, 60, 140], 100, 0.0)])     # no suppression
def test_detect_global_list_path(oracle, box, max_per, min_dist):
    """Targets whose candidates do not fit shared memory (or that ask for > 1024 corners) take the second candidate pass
    into the global list and the block-wide arg-max rounds; results as goodFeaturesToTrack's restatement."""
    w, h = 768, 576
    f = synth.Camera(w, h, 10, cam=8).frame(0)
    with klt.KLTContext(w, h, levels=2, max_targets=64) as c:
        c.push_frame(f)
        pts, counts = c.detect([box], max_per, 0.01, min_dist)
        eig = c.read_min_eig()
    want = oracle.detect(f, np.array(box, np.float64), max_per, 0.01, min_dist, eig=eig)
    assert counts[0] == len(want) and np.array_equal(pts[0, :counts[0]], want)


def test_detect_mixed_roi_sizes(oracle):
    """C3-shaped targets (1080p, 256 corners per target): ROIs smaller and larger than the shared-memory tile, clipped by
    every frame border, and one outside the frame, in one launch."""
    w, h = 1920, 1080
    f = synth.Camera(w, h, 12, cam=4).frame(0)
    boxes = np.array([[300.0, 200.0, 60.0, 150.0],        # 9 000 px
                      [700.0, 100.0, 137.0, 330.0],       # 45 000 px
                      [-20.0, 700.0, 260.0, 400.0],       # clipped left / bottom
                      [1700.0, -30.0, 400.0, 123.0],      # clipped right / top
                      [1000.0, 500.0, 128.0, 128.0],      # 16 900 px with the ring: just over the 64 KB tile
                      [2500.0, 300.0, 50.0, 50.0]])       # outside the frame
    with klt.KLTContext(w, h, levels=2, max_targets=64) as c:
        c.push_frame(f)
        pts, counts = c.detect(boxes, 256, 0.01, 3.0)
        eig = c.read_min_eig()
    for t, b in enumerate(boxes):
        want = oracle.detect(f, b, 256, 0.01, 3.0, eig=eig)
        assert counts[t] == len(want) and np.array_equal(pts[t, :counts[t]], want), t
    assert counts[5] == 0 and counts[1] == 256


def test_detect_full_rig_vs_oracle(oracle):
    w, h, ntgt = 768, 576, 40
    cam = synth.Camera(w, h, ntgt, cam=4)
    f = cam.frame(0)
    with klt.KLTContext(w, h, levels=2) as c:
        c.push_frame(f)
        eig = c.read_min_eig()
        assert np.array_equal(eig, oracle.min_eig_map(f))
        boxes = cam.boxes(0)
        pts, counts = c.detect(boxes, 100, 0.01, 3.0)
        for t in range(ntgt):
            want = oracle.detect(f, boxes[t], 100, 0.01, 3.0, eig=eig)
            assert np.array_equal(pts[t, :counts[t]], want), t
        pts, counts = c.detect([[0, 0, w, h]], 1000, 0.001, 1.0)                   # large candidate list (global scratch)
        want = oracle.detect(f, [0, 0, w, h], 1000, 0.001, 1.0, eig=eig)
        assert np.array_equal(pts[0, :counts[0]], want)
        pts, counts = c.detect(np.zeros((0, 4)), 10)
        assert counts.shape == (0,)


def test_overlapped_batched_detect_matches_plain(oracle):
    """PSNW_MEM_OVERLAP: detect on the auxiliary stream while a track launch runs on the main stream; results must
    equal the plain call, for several cameras in one batched call."""
    import ctypes as C
    import torch
    w, h, ncam, ntgt, per = 384, 288, 3, 10, 40
    cams = [synth.Camera(w, h, ntgt, cam=20 + c) for c in range(ncam)]
    with klt.KLTContext(w, h, n_cameras=ncam, levels=3, fb_thresh=1.0) as c:
        f0 = [cm.frame(0) for cm in cams]; f1 = [cm.frame(1) for cm in cams]
        c.push_frames(f0); c.push_frames(f1)
        plain = [c.detect(cams[k].boxes(1), per, 0.01, 3.0, cam=k) for k in range(ncam)]
        boxes = [np.ascontiguousarray(cams[k].boxes(1)) for k in range(ncam)]
        d_pts = torch.zeros((ncam, ntgt, per, 2), dtype=torch.float32, device="cuda")
        d_cnt = torch.zeros((ncam, ntgt), dtype=torch.int32, device="cuda")
        jobs = (klt.DetectJob * ncam)()
        for k in range(ncam):
            jobs[k].cam, jobs[k].slot, jobs[k].boxes, jobs[k].n_targets = k, 0, boxes[k].ctypes.data, ntgt
            jobs[k].pts_xy, jobs[k].counts = d_pts[k].data_ptr(), d_cnt[k].data_ptr()
        torch.cuda.synchronize()
        c._ck(c.lib.psnw_klt_detect_batch(c._h, jobs, ncam, per, 0.01, 3.0, klt.MEM_DEVICE | klt.MEM_OVERLAP))
        pts = np.random.default_rng(0).uniform(20, 250, size=(2000, 2)).astype(np.float32)
        c.track(pts, cam=0)                                   # main-stream work in flight next to the detect
        c.sync()                                              # joins the auxiliary stream
        g_pts, g_cnt = d_pts.cpu().numpy(), d_cnt.cpu().numpy()
        for k in range(ncam):
            assert np.array_equal(g_cnt[k], plain[k][1])
            assert np.array_equal(g_pts[k], plain[k][0])
            assert (g_pts[k][0, g_cnt[k][0]:] == -1e6).all()  # unused slots carry the sentinel


def test_points_in_boxes(ctx):
    """Next row f-3: per (detection, tracker) contain counts, PSN_Rect::contain semantics in double."""
    rng = np.random.default_rng(9)
    nt, nd = 37, 23
    ns = rng.integers(0, 120, size=nt); ns[0] = 0
    off = np.concatenate([[0], np.cumsum(ns)]).astype(np.int32)
    pts = rng.uniform(-20, 800, size=(off[-1], 2)).astype(np.float32)
    boxes = np.stack([rng.uniform(-10, 700, nd), rng.uniform(-10, 500, nd), rng.uniform(5, 200, nd), rng.uniform(5, 300, nd)], 1)
    pts[:5] = [boxes[0, 0], boxes[0, 1]]                           # exactly on the inclusive edge
    pts[5:9] = np.float32([boxes[0, 0] + boxes[0, 2], boxes[0, 1]])  # near the exclusive edge (float32-rounded)
    got = ctx.points_in_boxes(pts, off, boxes)
    want = np.zeros((nd, nt), np.int32)
    for d, (x, y, w, h) in enumerate(boxes):
        px, py = pts[:, 0].astype(np.float64), pts[:, 1].astype(np.float64)
        inside = (px >= x) & (px < x + w) & (py >= y) & (py < y + h)
        for t in range(nt):
            want[d, t] = inside[off[t]:off[t + 1]].sum()
    assert np.array_equal(got, want)


def _medianflow_case(rng, n, n_valid_frac=1.0, dup=False, quant=False):
    pre = rng.uniform(20, 400, size=(n, 2)).astype(np.float32)
    s, t = rng.uniform(0.9, 1.1), rng.uniform(-6, 6, size=2)
    c = pre.mean(axis=0) if n else np.zeros(2)
    cur = ((pre - c) * s + c + t + rng.normal(0, 0.3, size=(n, 2))).astype(np.float32)
    if n:
        out = rng.random(n) < 0.15                                   # gross outliers
        cur[out] += rng.uniform(-40, 40, size=(int(out.sum()), 2)).astype(np.float32)
    if dup and n > 3:
        pre[1] = pre[0]; pre[3] = pre[2]                             # coincident source points: pairs without a ratio
    if quant:
        pre, cur = np.round(pre), np.round(cur)                      # many equal displacements and ratios: ties
    st = (rng.random(n) < n_valid_frac).astype(np.uint8)
    return pre.astype(np.float32), cur.astype(np.float32), st


def test_medianflow_estimator_equals_oracle(oracle):
    """PSNW_BOX_MEDIANFLOW (north_star's median-displacement / scale reduction; not the reference's estimator): the CUDA
    kernel against its C definition, bit for bit -- odd and even counts, empty and single-point targets, more than 128
    points (strided pair subset), the 4096-point capacity, a status filter, coincident points, ties."""
    rng = np.random.default_rng(21)
    sizes = [0, 1, 2, 3, 4, 7, 50, 100, 128, 129, 257, 1000, 4096]
    cases = []
    for k, n in enumerate(sizes):
        cases.append(_medianflow_case(rng, n, n_valid_frac=1.0 if k % 3 else 0.7, dup=(k % 4 == 1), quant=(k % 5 == 2)))
    cases.append(_medianflow_case(rng, 60, n_valid_frac=0.0))         # no valid point at all
    boxes = np.stack([np.array([rng.uniform(0, 300), rng.uniform(0, 300), rng.uniform(20, 80), rng.uniform(40, 160)]) for _ in cases])
    pre = np.concatenate([c[0] for c in cases]); cur = np.concatenate([c[1] for c in cases]); st = np.concatenate([c[2] for c in cases])
    off = np.concatenate([[0], np.cumsum([len(c[0]) for c in cases])]).astype(np.int32)
    with klt.KLTContext(64, 64, levels=2, max_points=1 << 14, max_targets=64) as ctx:
        for status in (st, None):
            got_b, got_m, got_c = ctx.box_update(boxes, pre, cur, off, status=status, mode=klt.BOX_MEDIANFLOW)
            for t, c in enumerate(cases):
                a, b = off[t], off[t + 1]
                want_b, want_m = oracle.box_medianflow(boxes[t], c[0], c[1], None if status is None else c[2])
                assert np.array_equal(got_b[t], want_b), (t, len(c[0]), got_b[t], want_b)
                assert np.array_equal(got_m[a:b], want_m), (t, len(c[0]))
                assert got_c[t] == int(want_m.sum())
