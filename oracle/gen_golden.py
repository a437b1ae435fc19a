"""Generates tests/golden/*.npz -- small input/output vectors from the only runnable implementation of the
path's third-party arithmetic (python cv2 4.13) plus a literal pure-Python transcription of the
reference's own LocalSearchKLT.  TEST INFRASTRUCTURE; run from the repo root:

    python -m oracle.gen_golden

The fixtures pin oracle/klt_oracle.c on CPU (tests/test_golden.py) and the CUDA path on the GPU box, where
/root/reference does not exist (tests/test_gpu_*.py).  cv2 is not imported by the tests that consume them.
"""
import os

import cv2
import numpy as np

from mcmtt_opticalflow_b200 import synth

from . import cv2_ref

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def local_search_klt_py(box, pre, cur):
    """Line-by-line transcription of psn_where/PSNWhere_Tracker2D.cpp:455-554 (std::sort and all)."""
    n = len(pre)
    mv, midx, dxs, dys = [], [], [], []
    for i in range(n):
        fx = np.float32(cur[i][0]) - np.float32(pre[i][0])       # cv::Point2f operator- (:479)
        fy = np.float32(cur[i][1]) - np.float32(pre[i][1])
        x, y = float(fx), float(fy)
        if np.sqrt(x * x + y * y) < 0.1 * 1.0:                   # :480-484
            continue
        mv.append((x, y)); midx.append(i); dxs.append(x); dys.append(y)
    if len(mv) < n * 0.5:                                        # :493-496
        return list(box), []
    dxs.sort(); dys.sort()                                       # :498-499
    win = box[2] * 0.2 * 1.0                                     # :502
    max_nx = max_ny = 0
    ex = ey = 0.0
    for k in range(len(mv)):                                     # :505-537
        nx = sum(1 for j in range(len(mv)) if abs(dxs[k] - dxs[j]) < win)
        ny = sum(1 for j in range(len(mv)) if abs(dys[k] - dys[j]) < win)
        if max_nx < nx:
            ex, max_nx = dxs[k], nx
        if max_ny < ny:
            ey, max_ny = dys[k], ny
    inl = [midx[k] for k in range(len(mv)) if np.sqrt((mv[k][0] - ex) ** 2 + (mv[k][1] - ey) ** 2) < win]   # :540-546
    return [box[0] + ex, box[1] + ey, box[2], box[3]], inl      # :549-553


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20261018)

    # ---- frames: 256x192 crops of the synthetic rig (textured + low-texture variant) -------------------------
    cam = synth.Camera(768, 576, 20, cam=3, low_texture=True)
    f0, f1 = cam.frame(0), cam.frame(1)
    y0, x0 = 180, 250
    a, b = f0[y0:y0 + 192, x0:x0 + 256].copy(), f1[y0:y0 + 192, x0:x0 + 256].copy()
    odd_a, odd_b = a[:191, :253].copy(), b[:191, :253].copy()

    d = {"a": a, "b": b}
    # pyramid + Scharr of `a` and of an odd-sized crop, straight from cv2
    for name, img in (("a", a), ("odd", odd_a)):
        cur = img
        for l in range(4):
            if l:
                cur = cv2.pyrDown(cur, borderType=cv2.BORDER_REFLECT_101)
            d["pyr_%s_%d" % (name, l)] = cur
            d["dx_%s_%d" % (name, l)] = cv2.Scharr(cur, cv2.CV_16S, 1, 0, borderType=cv2.BORDER_REFLECT_101)
            d["dy_%s_%d" % (name, l)] = cv2.Scharr(cur, cv2.CV_16S, 0, 1, borderType=cv2.BORDER_REFLECT_101)

    # LK: corners + uniformly random points incl. outside the frame
    corners = cv2.goodFeaturesToTrack(a, 300, 0.01, 3).reshape(-1, 2)
    rnd = rng.uniform(-25, 280, size=(200, 2)).astype(np.float32)
    pts = np.concatenate([corners, rnd]).astype(np.float32)
    d["lk_pts"] = pts
    cfgs = [(21, 21, 3), (21, 21, 2), (21, 21, 0), (15, 31, 3), (41, 41, 3), (40, 100, 3), (9, 9, 4)]
    d["lk_cfgs"] = np.array(cfgs, np.int32)
    for i, (ww, wh, ml) in enumerate(cfgs):
        o, st, er = cv2_ref.lk(a, b, pts, (ww, wh), ml)
        d["lk_out_%d" % i], d["lk_st_%d" % i], d["lk_err_%d" % i] = o, st, er
    o, st, er, fb = cv2_ref.track_fb(a, b, pts, (21, 21), 3, fb_thresh=1.0)
    d["fb_out"], d["fb_st"], d["fb_err"] = o, st, fb
    # odd-size pair
    o, st, er = cv2_ref.lk(odd_a, odd_b, pts, (21, 21), 3)
    d["odd_b"] = odd_b
    d["lk_odd_out"], d["lk_odd_st"] = o, st

    # Shi-Tomasi: map + selections under box masks
    d["eig_a"] = cv2.cornerMinEigenVal(a, 3, ksize=3)
    boxes = np.array([[20, 10, 40, 96], [100.5, 30.2, 55, 130], [-10, -5, 60, 120], [220, 100, 50, 120],
                      [0, 0, 256, 192], [60, 60, 3, 3]], np.float64)
    d["gftt_boxes"] = boxes
    sel = [(100, 0.01, 3.0), (256, 0.05, 5.0), (30, 0.01, 0.0)]
    d["gftt_cfgs"] = np.array(sel, np.float64)
    for bi, box in enumerate(boxes):
        for si, (mc, q, md) in enumerate(sel):
            d["gftt_%d_%d" % (bi, si)] = cv2_ref.gftt(a, box, int(mc), q, md)
    # exact-tie image
    tie = np.zeros((64, 64), np.uint8)
    for (x, y) in [(10, 10), (30, 10), (50, 10), (10, 40), (30, 40)]:
        tie[y:y + 6, x:x + 6] = 200
    d["tie_img"] = tie
    d["tie_out"] = cv2_ref.gftt(tie, [0, 0, 200, 200], 50, 0.01, 3.0)

    # LocalSearchKLT cases from the literal transcription
    cases = []
    for c in range(40):
        n = int(rng.integers(0, 120))
        pre = rng.uniform(0, 200, size=(n, 2)).astype(np.float32)
        mode = c % 5
        shift = rng.uniform(-6, 6, size=2)
        cur = pre + shift + rng.normal(0, 0.3, size=(n, 2))
        if mode == 1:                                   # mostly static
            cur = pre + rng.normal(0, 0.02, size=(n, 2))
        elif mode == 2 and n:                           # outliers
            k = n // 3
            cur[:k] = pre[:k] + rng.uniform(-30, 30, size=(k, 2))
        elif mode == 3 and n:                           # quantised displacements -> ties in the neighbour counts
            cur = pre + np.round(rng.uniform(-3, 3, size=(n, 2)))
        elif mode == 4 and n:                           # exactly half static
            cur[: n // 2] = pre[: n // 2]
        cur = cur.astype(np.float32)
        box = [float(rng.uniform(0, 100)), float(rng.uniform(0, 100)), float(rng.uniform(8, 60)), float(rng.uniform(20, 140))]
        nb, inl = local_search_klt_py(box, pre, cur)
        cases.append((box, pre, cur, nb, inl))
    d["ls_n"] = np.array([len(c[1]) for c in cases], np.int32)
    d["ls_box"] = np.array([c[0] for c in cases], np.float64)
    d["ls_pre"] = np.concatenate([c[1] for c in cases]).astype(np.float32)
    d["ls_cur"] = np.concatenate([c[2] for c in cases]).astype(np.float32)
    d["ls_box_out"] = np.array([c[3] for c in cases], np.float64)
    d["ls_inl_n"] = np.array([len(c[4]) for c in cases], np.int32)
    d["ls_inl"] = np.array([i for c in cases for i in c[4]], np.int32)

    path = os.path.join(OUT, "klt_golden.npz")
    np.savez_compressed(path, **d)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB;", "cv2", cv2.__version__)


if __name__ == "__main__":
    main()
