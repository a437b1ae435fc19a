"""CPU tier: pins oracle/klt_oracle.c against cv2 4.13 (the runnable build of the library the reference links)
on fresh synthetic data.  cv2 is test infrastructure here, exactly like the oracle."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from mcmtt_opticalflow_b200 import synth
from oracle import cv2_ref


@pytest.fixture(scope="module")
def frames():
    cam = synth.Camera(384, 288, 8, cam=5, low_texture=True)
    return cam, cam.frame(0), cam.frame(1)


@pytest.mark.parametrize("shape", [(288, 384), (287, 383), (33, 129), (16, 16), (75, 100)])
def test_pyrdown_scharr_bit_exact(oracle, frames, shape):
    img = frames[1][:shape[0], :shape[1]].copy()
    assert np.array_equal(oracle.pyr_down(img), cv2.pyrDown(img, borderType=cv2.BORDER_REFLECT_101))
    s = oracle.scharr(img)
    assert np.array_equal(s[..., 0], cv2.Scharr(img, cv2.CV_16S, 1, 0, borderType=cv2.BORDER_REFLECT_101))
    assert np.array_equal(s[..., 1], cv2.Scharr(img, cv2.CV_16S, 0, 1, borderType=cv2.BORDER_REFLECT_101))


@pytest.mark.parametrize("win,ml", [((21, 21), 3), ((21, 21), 2), ((21, 21), 0), ((15, 31), 3), ((41, 41), 3),
                                    ((40, 100), 3), ((9, 9), 4)])
@pytest.mark.parametrize("exact", [False, True])
def test_lk_matches_cv2(oracle, frames, win, ml, exact):
    cam, f0, f1 = frames
    rng = np.random.default_rng(7)
    pts, _ = synth.seeds_in_boxes(cam.boxes(0), 40, 384, 288, rng)
    pts = np.concatenate([pts, rng.uniform(-30, 420, size=(150, 2)).astype(np.float32)])
    pa, pb = oracle.build_pyramid(f0, ml), oracle.build_pyramid(f1, ml)
    o, st, er = oracle.lk_track(pa, pb, pts, win, ml, exact_sums=exact)
    o2, st2, er2 = cv2_ref.lk(f0, f1, pts, win, ml)
    assert (st == st2).mean() >= 0.995
    both = (st == 1) & (st2 == 1)
    d = np.linalg.norm(o - o2, axis=1)
    assert (d[both] <= 0.05).mean() >= 0.99
    assert np.median(d[both]) < 1e-3
    assert np.abs(er - er2)[both].max() < 0.1


def test_level_truncation_rule(oracle):
    # 768x576, maxLevel 3 -> win (40,100),(72,72),(96,96) keep 2; win 21/40 keep 3
    assert oracle.effective_max_level(768, 576, (40, 100), 3) == 2
    assert oracle.effective_max_level(768, 576, (72, 72), 3) == 2
    assert oracle.effective_max_level(768, 576, (96, 96), 3) == 2
    assert oracle.effective_max_level(768, 576, (21, 21), 3) == 3
    assert oracle.effective_max_level(768, 576, (40, 40), 3) == 3


def test_fb_matches_cv2(oracle, frames):
    cam, f0, f1 = frames
    rng = np.random.default_rng(11)
    pts, _ = synth.seeds_in_boxes(cam.boxes(0), 60, 384, 288, rng)
    pa, pb = oracle.build_pyramid(f0, 3), oracle.build_pyramid(f1, 3)
    o, st, er, fb = oracle.track_fb(pa, pb, pts)
    o2, st2, er2, fb2 = cv2_ref.track_fb(f0, f1, pts)
    assert (st == st2).mean() >= 0.995
    ok = (st == 1) & (st2 == 1)
    assert (np.linalg.norm(o - o2, axis=1)[ok] <= 0.05).mean() >= 0.99


def test_min_eig_map_and_gftt(oracle, frames):
    cam, f0, _ = frames
    e, e2 = oracle.min_eig_map(f0), cv2.cornerMinEigenVal(f0, 3, ksize=3)
    assert np.abs(e - e2).max() <= 2e-6 * e2.max()
    for box in list(cam.boxes(0)) + [np.array([-10., -5., 60, 120]), np.array([350., 250., 50, 120])]:
        for mc, q, md in [(100, 0.01, 3.0), (256, 0.05, 5.0), (30, 0.01, 0.0), (100, 0.01, 2.5)]:
            a = oracle.detect(f0, box, mc, q, md, eig=e2)      # cv2's own map: isolates the selection logic
            b = cv2_ref.gftt(f0, box, mc, q, md)
            assert a.shape == b.shape and np.array_equal(a, b)
            a2 = oracle.detect(f0, box, mc, q, md)              # own map: float rounding may reorder near-ties
            sa, sb = set(map(tuple, a2)), set(map(tuple, b))
            assert len(sa ^ sb) <= max(2, len(sb) // 50)


def test_gftt_tie_order(oracle):
    img = np.zeros((64, 64), np.uint8)
    for (x, y) in [(10, 10), (30, 10), (50, 10), (10, 40), (30, 40)]:
        img[y:y + 6, x:x + 6] = 200
    c = cv2_ref.gftt(img, [0, 0, 200, 200], 50, 0.01, 3.0)
    assert np.array_equal(oracle.detect(img, [0, 0, 200, 200], 50, 0.01, 3.0), c)


def test_flow_majority_rule_hand_cases(oracle):
    """The run-length rule of Tracker2D.cpp:991-1013 on lists small enough to walk by hand."""
    det = np.array([[0.0, 0.0, 100.0, 100.0]])
    box = np.array([[10.0, 10.0, 20.0, 20.0]])

    def run(sizes, ok=None):
        n = len(sizes)
        off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        pts = np.full((off[-1], 2), 50.0, np.float32)                 # every point inside the detection box
        return oracle.flow_majority(pts, off, np.repeat(box, n, 0), np.ones(n, np.uint8) if ok is None else ok, det, 0.5)

    c, w = run([7]);            assert c.tolist() == [[7]] and w.tolist() == [0]            # sole tracker: 7 > 3.5
    c, w = run([6, 4]);         assert w.tolist() == [0]                                     # first run 6 > 5
    c, w = run([5, 5]);         assert w.tolist() == [1]                                     # 5 > 5 fails: the row is wiped
    c, w = run([2, 9]);         assert w.tolist() == [1]                                     # the last run is never compared: 2 > 5.5 fails
    c, w = run([2, 9, 1]);      assert w.tolist() == [0]                                     # a middle run counts 9 - 1 = 8 > 6
    c, w = run([3, 0, 5], ok=np.array([1, 1, 0], np.uint8)); assert c.tolist() == [[3, 0, 0]] and w.tolist() == [0]   # gated tracker: sole run
