"""CPU tier: oracle/klt_oracle.c against the committed fixtures (tests/golden/klt_golden.npz, made by
oracle/gen_golden.py from cv2 4.13 and a literal transcription of the reference's LocalSearchKLT)."""
import numpy as np
import pytest


def test_pyramid_and_scharr(oracle, golden):
    for name in ("a", "odd"):
        cur = golden["pyr_%s_0" % name]
        for l in range(4):
            if l:
                cur = oracle.pyr_down(cur)
            assert np.array_equal(cur, golden["pyr_%s_%d" % (name, l)])
            s = oracle.scharr(cur)
            assert np.array_equal(s[..., 0], golden["dx_%s_%d" % (name, l)])
            assert np.array_equal(s[..., 1], golden["dy_%s_%d" % (name, l)])


@pytest.mark.parametrize("exact", [False, True])
def test_lk(oracle, golden, exact):
    a, b, pts = golden["a"], golden["b"], golden["lk_pts"]
    for i, (ww, wh, ml) in enumerate(golden["lk_cfgs"]):
        pa, pb = oracle.build_pyramid(a, int(ml)), oracle.build_pyramid(b, int(ml))
        o, st, er = oracle.lk_track(pa, pb, pts, (int(ww), int(wh)), int(ml), exact_sums=exact)
        g_o, g_st, g_er = golden["lk_out_%d" % i], golden["lk_st_%d" % i], golden["lk_err_%d" % i]
        assert (st == g_st).mean() >= 0.995, (ww, wh, ml)
        ok = (st == 1) & (g_st == 1)
        d = np.linalg.norm(o - g_o, axis=1)
        assert (d[ok] <= 0.05).mean() >= 0.99, (ww, wh, ml)
        assert np.abs(er - g_er)[ok].max() < 0.1


def test_lk_odd_size(oracle, golden):
    a, b = golden["pyr_odd_0"], golden["odd_b"]
    pa, pb = oracle.build_pyramid(a, 3), oracle.build_pyramid(b, 3)
    o, st, _ = oracle.lk_track(pa, pb, golden["lk_pts"], (21, 21), 3)
    assert (st == golden["lk_odd_st"]).mean() >= 0.995
    ok = (st == 1) & (golden["lk_odd_st"] == 1)
    assert (np.linalg.norm(o - golden["lk_odd_out"], axis=1)[ok] <= 0.05).mean() >= 0.99


def test_fb(oracle, golden):
    a, b, pts = golden["a"], golden["b"], golden["lk_pts"]
    pa, pb = oracle.build_pyramid(a, 3), oracle.build_pyramid(b, 3)
    o, st, er, fb = oracle.track_fb(pa, pb, pts, fb_thresh=1.0)
    assert (st == golden["fb_st"]).mean() >= 0.995
    ok = (st == 1) & (golden["fb_st"] == 1)
    assert (np.linalg.norm(o - golden["fb_out"], axis=1)[ok] <= 0.05).mean() >= 0.99
    assert np.abs(fb - golden["fb_err"])[ok].max() < 0.05


def test_min_eig_and_selection(oracle, golden):
    a = golden["a"]
    e = oracle.min_eig_map(a)
    assert np.abs(e - golden["eig_a"]).max() <= 2e-6 * golden["eig_a"].max()
    for bi, box in enumerate(golden["gftt_boxes"]):
        for si, (mc, q, md) in enumerate(golden["gftt_cfgs"]):
            got = oracle.detect(a, box, int(mc), float(q), float(md), eig=golden["eig_a"])
            want = golden["gftt_%d_%d" % (bi, si)]
            assert got.shape == want.shape and np.array_equal(got, want), (bi, si)
    assert np.array_equal(oracle.detect(golden["tie_img"], [0, 0, 200, 200], 50, 0.01, 3.0), golden["tie_out"])


def test_local_search_klt(oracle, golden):
    off = np.concatenate([[0], np.cumsum(golden["ls_n"])])
    ioff = np.concatenate([[0], np.cumsum(golden["ls_inl_n"])])
    for c in range(len(golden["ls_n"])):
        pre, cur = golden["ls_pre"][off[c]:off[c + 1]], golden["ls_cur"][off[c]:off[c + 1]]
        box, inl = oracle.local_search_klt(golden["ls_box"][c], pre, cur)
        assert np.array_equal(box, golden["ls_box_out"][c]), c
        assert np.array_equal(inl, golden["ls_inl"][ioff[c]:ioff[c + 1]]), c


def test_local_search_klt_properties(oracle):
    rng = np.random.default_rng(3)
    pre = rng.uniform(0, 100, size=(80, 2)).astype(np.float32)
    cur = (pre + np.float32([2.5, -1.25])).astype(np.float32)
    box = np.array([10.0, 20.0, 40.0, 96.0])
    nb, inl = oracle.local_search_klt(box, pre, cur)
    assert np.allclose(nb, [12.5, 18.75, 40.0, 96.0], atol=1e-4) and len(inl) == 80      # pure translation
    assert np.array_equal(nb[2:], box[2:])                                                # w, h never change
    nb2, inl2 = oracle.local_search_klt(box, pre, pre)                                    # static shortcut
    assert np.array_equal(nb2, box) and len(inl2) == 0
    # fewer than half moving -> old box, empty inlier list (Tracker2D.cpp:493-496)
    cur3 = pre.copy(); cur3[:39] += 3
    nb3, inl3 = oracle.local_search_klt(box, pre, cur3)
    assert np.array_equal(nb3, box) and len(inl3) == 0
    # status filter == passing the filtered subset
    valid = (rng.random(80) < 0.6).astype(np.uint8)
    nb4, inl4 = oracle.local_search_klt(box, pre, cur, valid)
    nb5, inl5 = oracle.local_search_klt(box, pre[valid == 1], cur[valid == 1])
    assert np.array_equal(nb4, nb5) and np.array_equal(np.flatnonzero(valid)[inl5], inl4)
    # empty input
    nb6, inl6 = oracle.local_search_klt(box, np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32))
    assert np.array_equal(nb6, box) and len(inl6) == 0


def test_medianflow_definition_on_a_clean_motion():
    """The median-flow restatement (oracle/klt_oracle.c::psnwo_box_medianflow) recovers a pure translation + scale, equals
    numpy's medians on distinct values and keeps at least half of the points."""
    from oracle import psnw_oracle as po
    po.build()
    rng = np.random.default_rng(3)
    pre = rng.uniform(50, 300, size=(101, 2)).astype(np.float64)
    c, s, t = pre.mean(axis=0), 1.0625, np.array([3.5, -2.25])
    cur = (pre - c) * s + c + t
    pre32, cur32 = pre.astype(np.float32), cur.astype(np.float32)
    box = np.array([100.0, 80.0, 40.0, 90.0])
    out, mask = po.box_medianflow(box, pre32, cur32)
    d = cur32.astype(np.float64) - pre32.astype(np.float64)
    assert out[2] == box[2] * (out[2] / box[2]) and abs(out[2] / box[2] - s) < 1e-4 and abs(out[3] / box[3] - s) < 1e-4
    mdx, mdy = np.median(d[:, 0]), np.median(d[:, 1])
    sc = out[2] / box[2]
    assert abs(out[0] - (box[0] + mdx - 0.5 * (sc - 1) * box[2])) < 1e-9 and abs(out[1] - (box[1] + mdy - 0.5 * (sc - 1) * box[3])) < 1e-9
    assert mask.sum() >= 51
    out0, mask0 = po.box_medianflow(box, pre32, cur32, np.zeros(101, np.uint8))
    assert np.array_equal(out0, box) and mask0.sum() == 0
