"""GridFAST ("next" row f-1): the reference's real detector (psn_where/PSNWhere_Tracker2D.cpp:142, :735-757).
CPU tier: the C restatement against cv2 4.13's FastFeatureDetector run per grid cell.  GPU tier: the CUDA path
against the C restatement, bit-exact."""
import numpy as np
import pytest

from mcmtt_opticalflow_b200 import synth


@pytest.fixture(scope="module")
def scene():
    cam = synth.Camera(768, 576, 20, cam=2, low_texture=True)
    return cam, cam.frame(0)


def test_oracle_matches_cv2_fast_per_cell(oracle, scene):
    pytest.importorskip("cv2")
    from oracle import cv2_ref
    cam, f = scene
    fm = oracle.gridfast_map(f)
    pts, resp = cv2_ref.gridfast(f, [0, 0, 768, 576], max_total=10 ** 9)           # every keypoint of every cell
    mine = np.argwhere(fm > 0)[:, ::-1].astype(np.float32)
    mine = mine[np.lexsort((mine[:, 0], mine[:, 1]))]
    assert len(mine) > 1000 and np.array_equal(pts, mine)
    assert np.array_equal(fm[mine[:, 1].astype(int), mine[:, 0].astype(int)].astype(np.float32), resp)
    for box in list(cam.boxes(0)) + [np.array([-10., -5., 60, 120]), np.array([700., 480., 80, 120]), np.array([0., 0., 768, 576])]:
        a = oracle.gridfast_detect(f, box, max_out=10 ** 6, fmap=fm)
        b, _ = cv2_ref.gridfast(f, box)
        assert np.array_equal(a, b)
    odd = f[:301, :403].copy()                                                       # cell edges not multiples of 4
    pts, _ = cv2_ref.gridfast(odd, [0, 0, 403, 301], max_total=10 ** 9, grid=(3, 5))
    fm = oracle.gridfast_map(odd, grid=(3, 5))
    mine = np.argwhere(fm > 0)[:, ::-1].astype(np.float32)
    assert np.array_equal(pts, mine[np.lexsort((mine[:, 0], mine[:, 1]))])


def test_oracle_cap_keeps_strongest(oracle, scene):
    cam, f = scene
    fm = oracle.gridfast_map(f)
    full = oracle.gridfast_detect(f, [0, 0, 768, 576], max_out=10 ** 6, fmap=fm)
    top = oracle.gridfast_detect(f, [0, 0, 768, 576], max_out=100, fmap=fm)
    assert len(top) == 100
    s_full = np.sort(fm[full[:, 1].astype(int), full[:, 0].astype(int)])[::-1]
    s_top = fm[top[:, 1].astype(int), top[:, 0].astype(int)]
    assert s_top.min() >= s_full[99] and set(map(tuple, top)) <= set(map(tuple, full))


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,grid", [(768, 576, (4, 4)), (403, 301, (3, 5)), (130, 70, (4, 4))])
def test_gpu_gridfast_bit_exact(oracle, w, h, grid):
    from mcmtt_opticalflow_b200 import klt
    cam = synth.Camera(768, 576, 20, cam=2, low_texture=True)
    f = np.ascontiguousarray(cam.frame(0)[:h, :w])
    boxes = np.concatenate([cam.boxes(0)[:12] * (w / 768.0), [[-10, -5, 60, 120], [w - 60, h - 90, 80, 120], [0, 0, w, h]]])
    with klt.KLTContext(w, h, levels=2) as c:
        c.push_frame(f)
        for cap in (100, 1000, 7):
            pts, counts = c.gridfast_detect(boxes, cap, 10, grid, 1000)
            fm = c.read_fast_map()
            assert np.array_equal(fm, oracle.gridfast_map(f, 10, grid))
            for t, box in enumerate(boxes):
                want = oracle.gridfast_detect(f, box, cap, 10, grid, 1000, fmap=fm)
                assert counts[t] == len(want) and np.array_equal(pts[t, :counts[t]], want), (t, cap)
                assert (pts[t, counts[t]:] == -1e6).all()
        pts, counts = c.gridfast_detect(boxes[:2], 50, 25, grid, 32)                  # other threshold / per-cell cap
        fm = c.read_fast_map()
        assert np.array_equal(fm, oracle.gridfast_map(f, 25, grid))
        for t in range(2):
            want = oracle.gridfast_detect(f, boxes[t], 50, 25, grid, 32, fmap=fm)
            assert np.array_equal(pts[t, :counts[t]], want)
