"""ctypes front-end of the CPU oracle (oracle/klt_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module; nothing in mcmtt_opticalflow_b200/ does.  Each function names the reference site it
restates (paths relative to /root/reference):

  build_pyramid / lk_track   cv::calcOpticalFlowPyrLK as called at psn_where/PSNWhere_Tracker2D.cpp:776-782, :871-877
  track_fb                   north_star forward-backward rule
  local_search_klt           psn_where/PSNWhere_Tracker2D.cpp:452-554
  detect                     call site psn_where/PSNWhere_Tracker2D.cpp:735-757 with north_star's Shi-Tomasi detector
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpsnw_oracle.so")


def build(force=False):
    """Compile the C restatement (gcc, seconds)."""
    src = os.path.join(_HERE, "klt_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libpsnw_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _Level(C.Structure):
    _fields_ = [("img", C.c_void_p), ("dxy", C.c_void_p), ("w", C.c_int), ("h", C.c_int),
                ("img_stride", C.c_int), ("dxy_stride_px", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.psnwo_effective_max_level.restype = C.c_int
        _lib.psnwo_local_search_klt.restype = C.c_int
        _lib.psnwo_gftt_select.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def pyr_down(img):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.empty(((h + 1) // 2, (w + 1) // 2), np.uint8)
    lib().psnwo_pyr_down(_p(img), w, h, img.strides[0], _p(out), out.strides[0])
    return out


def scharr(img):
    """-> int16 [h, w, 2] interleaved (dx, dy)."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.empty((h, w, 2), np.int16)
    lib().psnwo_scharr(_p(img), w, h, img.strides[0], _p(out), w)
    return out


def effective_max_level(w, h, win, max_level):
    return lib().psnwo_effective_max_level(w, h, int(win[0]), int(win[1]), int(max_level))


def build_pyramid(img, max_level):
    """All max_level+1 levels (u8, dxy); the per-window truncation is applied by lk_track."""
    levels = []
    cur = np.ascontiguousarray(img, np.uint8)
    for l in range(max_level + 1):
        if l > 0:
            cur = pyr_down(cur)
        levels.append((cur, scharr(cur)))
    return levels


def _levels_array(pyr):
    arr = (_Level * len(pyr))()
    for i, (u8, dxy) in enumerate(pyr):
        arr[i].img = u8.ctypes.data
        arr[i].dxy = dxy.ctypes.data
        arr[i].h, arr[i].w = u8.shape
        arr[i].img_stride = u8.strides[0]
        arr[i].dxy_stride_px = dxy.shape[1]
    return arr


def lk_track(pyr_from, pyr_to, pts, win=(21, 21), max_level=3, max_iters=30, epsilon=0.01,
             min_eig_thresh=1e-4, exact_sums=False):
    """One LK direction.  pts float32 [n,2] -> (out [n,2], status u8 [n], err f32 [n])."""
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    n = pts.shape[0]
    h, w = pyr_from[0][0].shape
    ml = effective_max_level(w, h, win, min(max_level, len(pyr_from) - 1))
    out = np.zeros((n, 2), np.float32)
    status = np.zeros(n, np.uint8)
    err = np.zeros(n, np.float32)
    la, lb = _levels_array(pyr_from), _levels_array(pyr_to)
    lib().psnwo_lk_track(la, lb, C.c_int(ml), _p(pts), C.c_int(n), C.c_int(int(win[0])), C.c_int(int(win[1])),
                         C.c_int(max_iters), C.c_double(epsilon), C.c_float(min_eig_thresh),
                         C.c_int(1 if exact_sums else 0), _p(out), _p(status), _p(err))
    return out, status, err


def track_fb(pyr_a, pyr_b, pts, win=(21, 21), max_level=3, max_iters=30, epsilon=0.01,
             min_eig_thresh=1e-4, fb_thresh=1.0, exact_sums=False):
    """Forward a->b, backward b->a, FB gate.  -> (fwd pts, status, err_fwd, fb_err)."""
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    p1, st_f, err_f = lk_track(pyr_a, pyr_b, pts, win, max_level, max_iters, epsilon, min_eig_thresh, exact_sums)
    p2, st_b, _ = lk_track(pyr_b, pyr_a, p1, win, max_level, max_iters, epsilon, min_eig_thresh, exact_sums)
    n = pts.shape[0]
    status = np.zeros(n, np.uint8)
    fb = np.zeros(n, np.float32)
    lib().psnwo_fb_check(_p(pts), _p(p2), _p(st_f), _p(st_b), C.c_int(n), C.c_float(fb_thresh), _p(status), _p(fb))
    return p1, status, err_f, fb


def local_search_klt(box, pre, cur, valid=None):
    """-> (new box float64[4], inlier indices int32[k])."""
    box = np.ascontiguousarray(box, np.float64).reshape(4)
    pre = np.ascontiguousarray(pre, np.float32).reshape(-1, 2)
    cur = np.ascontiguousarray(cur, np.float32).reshape(-1, 2)
    n = pre.shape[0]
    out = np.zeros(4, np.float64)
    idx = np.zeros(max(n, 1), np.int32)
    v = None if valid is None else np.ascontiguousarray(valid, np.uint8)
    k = lib().psnwo_local_search_klt(_p(box), _p(pre), _p(cur), _p(v) if v is not None else None,
                                     C.c_int(n), _p(out), _p(idx))
    return out, idx[:k].copy()


def box_medianflow(box, pre, cur, valid=None):
    """PSNW_BOX_MEDIANFLOW (include/psnw_klt.h) -> (new box float64[4], inlier mask uint8[n])."""
    box = np.ascontiguousarray(box, np.float64).reshape(4)
    pre = np.ascontiguousarray(pre, np.float32).reshape(-1, 2)
    cur = np.ascontiguousarray(cur, np.float32).reshape(-1, 2)
    n = pre.shape[0]
    out = np.zeros(4, np.float64)
    mask = np.zeros(max(n, 1), np.uint8)
    v = None if valid is None else np.ascontiguousarray(valid, np.uint8)
    f = lib().psnwo_box_medianflow
    f.restype = C.c_int
    f(_p(box), _p(pre), _p(cur), _p(v) if v is not None else None, C.c_int(n), _p(out), _p(mask))
    return out, mask[:n].copy()


def flow_majority(pts, tracker_offsets, tracker_boxes, tracker_ok, det_boxes, ratio=0.5):
    """Optical-flow majority check (Tracker2D.cpp:931-944, :981-1022) -> (counts int32 [n_det, n_trk], wipe uint8 [n_det])."""
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    off = np.ascontiguousarray(tracker_offsets, np.int32)
    tb = np.ascontiguousarray(tracker_boxes, np.float64).reshape(-1, 4)
    ok = np.ascontiguousarray(tracker_ok, np.uint8)
    db = np.ascontiguousarray(det_boxes, np.float64).reshape(-1, 4)
    nt, nd = len(off) - 1, len(db)
    counts, wipe = np.zeros((max(nd, 1), max(nt, 1)), np.int32), np.zeros(max(nd, 1), np.uint8)
    lib().psnwo_flow_majority(_p(pts), _p(off), C.c_int(nt), _p(tb), _p(ok), _p(db), C.c_int(nd), C.c_double(ratio), _p(counts), _p(wipe))
    return counts[:nd, :nt], wipe[:nd]


def min_eig_map(img):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.empty((h, w), np.float32)
    lib().psnwo_min_eig_map(_p(img), w, h, img.strides[0], _p(out), w)
    return out


def crop_roi(box, w, h):
    """PSN_Rect::cropWithSize(w,h).cv()  (psn_where/PSNWhere_Types.h:146-153,181; call site Tracker2D.cpp:735)."""
    x, y, bw, bh = [float(v) for v in box]
    nx, ny = max(0.0, x), max(0.0, y)
    nw, nh = min(w - nx - 1, bw), min(h - ny - 1, bh)
    return int(nx), int(ny), int(nx) + max(int(nw), 0), int(ny) + max(int(nh), 0)


def detect(img, box, max_corners=100, quality=0.01, min_dist=3.0, eig=None):
    """Shi-Tomasi corners inside one detection box -> float32 [k,2] strongest first."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    if eig is None:
        eig = min_eig_map(img)
    x0, y0, x1, y1 = crop_roi(box, w, h)
    out = np.zeros((max(max_corners, 1), 2), np.float32)
    k = lib().psnwo_gftt_select(_p(eig), w, h, eig.shape[1], x0, y0, x1, y1, C.c_int(max_corners),
                                C.c_float(quality), C.c_float(min_dist), _p(out))
    return out[:k].copy()


def gridfast_map(img, threshold=10, grid=(4, 4)):
    """NMS'd FAST-9/16 score map with GridAdaptedFeatureDetector cell semantics -> u8 [h, w]."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.zeros((h, w), np.uint8)
    lib().psnwo_gridfast_map(_p(img), w, h, img.strides[0], int(threshold), int(grid[0]), int(grid[1]), _p(out), w)
    return out


def gridfast_detect(img, box, max_out=100, threshold=10, grid=(4, 4), max_total=1000, fmap=None):
    """GridFAST keypoints inside one detection box (Tracker2D.cpp:735-757 without the random shuffle) -> f32 [k,2]."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    if fmap is None:
        fmap = gridfast_map(img, threshold, grid)
    x0, y0, x1, y1 = crop_roi(box, w, h)
    out = np.zeros((max(max_out, 1), 2), np.float32)
    lib().psnwo_gridfast_select.restype = C.c_int
    k = lib().psnwo_gridfast_select(_p(fmap), w, h, fmap.shape[1], int(grid[0]), int(grid[1]), int(max_total),
                                    x0, y0, x1, y1, int(max_out), _p(out))
    return out[:k].copy()
