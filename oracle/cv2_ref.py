"""cv2 (OpenCV 4.13, the only runnable build of the library the reference links) as the pin for
oracle/klt_oracle.c and as the CPU baseline.  TEST INFRASTRUCTURE ONLY -- same import rule as
psnw_oracle.py.

The reference calls OpenCV 2.4.6 at psn_where/PSNWhere_Tracker2D.cpp:741,776,871; the LK / pyrDown /
Scharr / min-eig algorithms are unchanged between 2.4.6 and 4.13 in everything this path uses
(SURVEY.md section 8c, Appendix A/D).
"""
import cv2
import numpy as np

from . import psnw_oracle as po


def lk(img_a, img_b, pts, win=(21, 21), max_level=3, max_iters=30, epsilon=0.01, min_eig_thresh=1e-4):
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 1, 2)
    if pts.shape[0] == 0:
        return np.zeros((0, 2), np.float32), np.zeros(0, np.uint8), np.zeros(0, np.float32)
    crit = (cv2.TERM_CRITERIA_COUNT | cv2.TERM_CRITERIA_EPS, max_iters, epsilon)
    out, st, err = cv2.calcOpticalFlowPyrLK(img_a, img_b, pts, None, winSize=(int(win[0]), int(win[1])),
                                            maxLevel=max_level, criteria=crit, flags=0,
                                            minEigThreshold=min_eig_thresh)
    return out.reshape(-1, 2), st.reshape(-1).astype(np.uint8), err.reshape(-1)


def track_fb(img_a, img_b, pts, win=(21, 21), max_level=3, max_iters=30, epsilon=0.01,
             min_eig_thresh=1e-4, fb_thresh=1.0):
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    p1, st_f, err_f = lk(img_a, img_b, pts, win, max_level, max_iters, epsilon, min_eig_thresh)
    p2, st_b, _ = lk(img_b, img_a, p1, win, max_level, max_iters, epsilon, min_eig_thresh)
    d = pts - p2
    fb = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32)
    fb[st_f == 0] = 0.0          # points the forward pass lost: no backward pass, fb_err defined as 0 (oracle/klt_oracle.c::psnwo_fb_check)
    ok = (st_f != 0) & (st_b != 0)
    if fb_thresh >= 0:
        ok &= fb <= np.float32(fb_thresh)
    return p1, ok.astype(np.uint8), err_f, fb


def gftt(img, box, max_corners=100, quality=0.01, min_dist=3.0):
    """cv2.goodFeaturesToTrack under the reference's box mask (Tracker2D.cpp:735-742)."""
    h, w = img.shape
    x0, y0, x1, y1 = po.crop_roi(box, w, h)
    mask = np.zeros((h, w), np.uint8)
    mask[y0:y1, x0:x1] = 255
    c = cv2.goodFeaturesToTrack(img, max_corners, quality, min_dist, mask=mask, blockSize=3, useHarrisDetector=False)
    return np.zeros((0, 2), np.float32) if c is None else c.reshape(-1, 2)


def box_iou(a, b):
    """IoU from PSN_Rect::overlappedArea / area (psn_where/PSNWhere_Types.h:158,171-178)."""
    ow = min(a[0] + a[2], b[0] + b[2]) - max(a[0], b[0])
    oh = min(a[1] + a[3], b[1] + b[3]) - max(a[1], b[1])
    if ow <= 0 or oh <= 0:
        return 0.0
    inter = ow * oh
    return inter / (a[2] * a[3] + b[2] * b[3] - inter)


def gridfast(img, box, threshold=10, grid=(4, 4), max_total=1000):
    """OpenCV 2.4.x GridAdaptedFeatureDetector(FAST(threshold, nms), max_total, grid) under the reference's box mask,
    rebuilt from cv2 4.13's FastFeatureDetector run on every cell's sub-image (the class itself left OpenCV in 3.0).
    -> (points float32 [k,2] in raster order, responses) of the keypoints that survive keepStrongest per cell."""
    h, w = img.shape
    x0, y0, x1, y1 = po.crop_roi(box, w, h)
    fast = cv2.FastFeatureDetector_create(threshold=threshold, nonmaxSuppression=True)
    per_cell = max_total // (grid[0] * grid[1])
    pts = []
    for i in range(grid[0]):
        r0, r1 = (i * h) // grid[0], ((i + 1) * h) // grid[0]
        for j in range(grid[1]):
            c0, c1 = (j * w) // grid[1], ((j + 1) * w) // grid[1]
            kps = fast.detect(np.ascontiguousarray(img[r0:r1, c0:c1]), None)
            cell = [(int(k.pt[0]) + c0, int(k.pt[1]) + r0, k.response) for k in kps]
            cell = [p for p in cell if x0 <= p[0] < x1 and y0 <= p[1] < y1]            # runByPixelsMask
            if len(cell) > per_cell:                                                   # retainBest keeps boundary ties
                thr = sorted((p[2] for p in cell), reverse=True)[per_cell - 1]
                cell = [p for p in cell if p[2] >= thr]
            pts += cell
    pts.sort(key=lambda p: (p[1], p[0]))
    return (np.array([[p[0], p[1]] for p in pts], np.float32).reshape(-1, 2), np.array([p[2] for p in pts], np.float32))
