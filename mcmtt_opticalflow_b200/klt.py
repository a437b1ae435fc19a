"""ctypes binding of the C ABI in include/psnw_klt.h -- the only way Python reaches the CUDA kernels.

The library is built in-tree (mcmtt_opticalflow_b200/libpsnw_klt.so, see csrc/Makefile).  If it is
missing, or no CUDA device is present, construction raises: there is no CPU fallback and nothing in
this package imports oracle/.

Reference mapping (paths relative to /root/reference/psn_where):
  KLTContext.push_frames   PSNWhere_Tracker2D.cpp:256-263 + ring :70,:311-316
  KLTContext.detect        PSNWhere_Tracker2D.cpp:735-757
  KLTContext.track         PSNWhere_Tracker2D.cpp:776-782 / :871-877 (cv::calcOpticalFlowPyrLK)
  KLTContext.box_update    PSNWhere_Tracker2D.cpp:452-554 (LocalSearchKLT)
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PSNW_KLT_LIB") or os.path.join(_HERE, "libpsnw_klt.so")

OK, EARG, ECUDA, ENOMEM, ESTATE = 0, -1, -2, -3, -4
MEM_HOST, MEM_PINNED, MEM_DEVICE = 0, 1, 2
MEM_OVERLAP = 0x100
PUSH_EIG = 0x200
BOX_REFERENCE = 0
BOX_MEDIANFLOW = 1
_ERR = {EARG: "PSNW_EARG", ECUDA: "PSNW_ECUDA", ENOMEM: "PSNW_ENOMEM", ESTATE: "PSNW_ESTATE"}


class PsnwError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s (%d): %s" % (_ERR.get(code, "PSNW_E?"), code, msg))
        self.code = code


class Params(C.Structure):
    _fields_ = [("device", C.c_int), ("n_cameras", C.c_int), ("width", C.c_int), ("height", C.c_int),
                ("levels", C.c_int), ("ring_slots", C.c_int), ("win_w", C.c_int), ("win_h", C.c_int),
                ("max_iters", C.c_int), ("epsilon", C.c_double), ("min_eig_thresh", C.c_float),
                ("fb_thresh", C.c_float), ("max_points", C.c_int), ("max_targets", C.c_int),
                ("stream", C.c_void_p), ("apron", C.c_int)]


class TrackJob(C.Structure):
    _fields_ = [("cam", C.c_int), ("slot_from", C.c_int), ("slot_to", C.c_int), ("n_pts", C.c_int),
                ("pts_xy", C.c_void_p), ("target_of_pt", C.c_void_p), ("n_targets", C.c_int),
                ("win_wh", C.c_void_p), ("out_xy", C.c_void_p), ("status", C.c_void_p), ("err", C.c_void_p),
                ("fb_err", C.c_void_p)]


class DetectJob(C.Structure):
    _fields_ = [("cam", C.c_int), ("slot", C.c_int), ("boxes", C.c_void_p), ("n_targets", C.c_int),
                ("pts_xy", C.c_void_p), ("counts", C.c_void_p)]


EXPORTS = [
    "psnw_klt_detect_batch", "psnw_join", "psnw_gridfast_detect", "psnw_read_fast_map", "psnw_points_in_boxes", "psnw_default_params", "psnw_create", "psnw_destroy", "psnw_push_frame", "psnw_push_frames",
    "psnw_push_frames_bgr", "psnw_frames_held", "psnw_klt_detect", "psnw_klt_track", "psnw_klt_track_batch",
    "psnw_box_update", "psnw_sync", "psnw_last_error", "psnw_get_stream", "psnw_level_info", "psnw_read_level",
    "psnw_read_level_apron", "psnw_klt_track_update", "psnw_get_params",
    "psnw_read_min_eig", "psnw_launch_count", "psnw_push_frame_bytes", "psnw_rebuild_pyramids", "psnw_levels_built",
    "psnw_rebuild_pyramids_slot", "psnw_lk_stats", "psnw_klt_backward_chain", "psnw_flow_majority",
]

_lib = None


def load():
    """Load libpsnw_klt.so; raises if the CUDA extension has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: build it with `make -C mcmtt_opticalflow_b200/csrc` "
                              "(or __graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        lib.psnw_last_error.restype = C.c_char_p
        lib.psnw_last_error.argtypes = [C.c_void_p]
        lib.psnw_get_stream.restype = C.c_void_p
        lib.psnw_get_stream.argtypes = [C.c_void_p]
        lib.psnw_launch_count.restype = C.c_uint64
        lib.psnw_launch_count.argtypes = [C.c_void_p]
        lib.psnw_push_frame_bytes.restype = C.c_uint64
        lib.psnw_push_frame_bytes.argtypes = [C.c_void_p]
        lib.psnw_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Params)]
        lib.psnw_destroy.argtypes = [C.c_void_p]
        lib.psnw_get_params.argtypes = [C.c_void_p, C.POINTER(Params)]
        lib.psnw_sync.argtypes = [C.c_void_p]
        lib.psnw_join.argtypes = [C.c_void_p]
        lib.psnw_frames_held.argtypes = [C.c_void_p, C.c_int]
        lib.psnw_levels_built.argtypes = [C.c_void_p]
        lib.psnw_push_frames.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_int]
        lib.psnw_push_frames_bgr.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32,
                                             C.c_int, C.c_int]
        lib.psnw_push_frame.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_uint32, C.c_int]
        lib.psnw_rebuild_pyramids.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.psnw_rebuild_pyramids_slot.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        lib.psnw_lk_stats.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.psnw_klt_backward_chain.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.psnw_flow_majority.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                           C.c_double, C.c_void_p, C.c_void_p, C.c_int]
        lib.psnw_klt_detect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float,
                                        C.c_float, C.c_void_p, C.c_void_p, C.c_int]
        lib.psnw_klt_detect_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int]
        lib.psnw_gridfast_detect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        lib.psnw_read_fast_map.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.psnw_points_in_boxes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        lib.psnw_klt_track.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                       C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.psnw_klt_track_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        if hasattr(lib, "psnw_klt_track_update"):
            lib.psnw_klt_track_update.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                                  C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                  C.c_void_p, C.c_int]
        lib.psnw_box_update.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.psnw_level_info.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.psnw_read_level.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        if hasattr(lib, "psnw_read_level_apron"):       # absent only from older bring-up builds named by PSNW_KLT_LIB
            lib.psnw_read_level_apron.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        lib.psnw_read_min_eig.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib = lib
    return _lib


def default_params():
    p = Params()
    load().psnw_default_params(C.byref(p))
    return p


def _ptr(a):
    """numpy array -> address; int -> itself (device pointer); None -> NULL."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return int(a)
    return a.ctypes.data


class KLTContext:
    """One GPU's shard of the camera rig: `n_cameras` cameras with a 4-deep device ring each."""

    def __init__(self, width, height, n_cameras=1, levels=4, win=(21, 21), ring_slots=4, max_iters=30,
                 epsilon=0.01, min_eig_thresh=1e-4, fb_thresh=1.0, max_points=1 << 16, max_targets=1 << 12,
                 device=0, stream=None, apron=0):
        self.lib = load()
        p = default_params()
        p.device, p.n_cameras, p.width, p.height = device, n_cameras, width, height
        p.levels, p.ring_slots, p.win_w, p.win_h = levels, ring_slots, int(win[0]), int(win[1])
        p.max_iters, p.epsilon, p.min_eig_thresh, p.fb_thresh = max_iters, epsilon, min_eig_thresh, fb_thresh
        p.max_points, p.max_targets = max_points, max_targets
        p.stream = stream
        p.apron = apron
        self.params = p
        self.width, self.height, self.n_cameras = width, height, n_cameras
        self.do_fb = fb_thresh >= 0
        self._h = C.c_void_p()
        rc = self.lib.psnw_create(C.byref(self._h), C.byref(p))
        if rc != OK:
            self._h = C.c_void_p()
            raise PsnwError(rc, "psnw_create failed (no CUDA device, bad parameters or out of memory)")
        self.levels = self.lib.psnw_levels_built(self._h)

    # -- plumbing ------------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != OK:
            raise PsnwError(rc, (self.lib.psnw_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.psnw_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def sync(self):
        self._ck(self.lib.psnw_sync(self._h))

    def join(self):
        self._ck(self.lib.psnw_join(self._h))

    @property
    def stream(self):
        return self.lib.psnw_get_stream(self._h)

    @property
    def launch_count(self):
        return int(self.lib.psnw_launch_count(self._h))

    @property
    def push_frame_bytes(self):
        return int(self.lib.psnw_push_frame_bytes(self._h))

    def frames_held(self, cam=0):
        return self.lib.psnw_frames_held(self._h, cam)

    def level_info(self, level):
        w, h, pitch = C.c_int(), C.c_int(), C.c_int()
        self._ck(self.lib.psnw_level_info(self._h, level, C.byref(w), C.byref(h), C.byref(pitch)))
        return w.value, h.value, pitch.value

    # -- frame intake --------------------------------------------------------------------------
    def push_frames(self, frames, cams=None, frame_idx=0, mem=MEM_HOST, pitch=None):
        """frames: list of u8 [h, w] arrays (or device/pinned addresses with `pitch`)."""
        n = len(frames)
        cams = np.arange(n, dtype=np.int32) if cams is None else np.ascontiguousarray(cams, np.int32)
        keep = []
        ptrs = (C.c_void_p * n)()
        for i, f in enumerate(frames):
            if isinstance(f, (int, np.integer)):
                ptrs[i] = int(f)
            else:
                f = np.ascontiguousarray(f, np.uint8)
                assert f.shape == (self.height, self.width), f.shape
                keep.append(f)
                ptrs[i] = f.ctypes.data
                pitch = f.strides[0]
        self._ck(self.lib.psnw_push_frames(self._h, n, cams.ctypes.data, ptrs, pitch or self.width, frame_idx, mem))

    def push_frame(self, frame, cam=0, frame_idx=0, mem=MEM_HOST):
        self.push_frames([frame], [cam], frame_idx, mem)

    def push_frames_bgr(self, frames, cams=None, frame_idx=0, luma_bits=15, mem=MEM_HOST):
        n = len(frames)
        cams = np.arange(n, dtype=np.int32) if cams is None else np.ascontiguousarray(cams, np.int32)
        fr = [np.ascontiguousarray(f, np.uint8) for f in frames]
        ptrs = (C.c_void_p * n)(*[f.ctypes.data for f in fr])
        self._ck(self.lib.psnw_push_frames_bgr(self._h, n, cams.ctypes.data, ptrs, fr[0].strides[0], frame_idx,
                                               luma_bits, mem))

    def rebuild_pyramids(self, cams, slot=0):
        cams = np.ascontiguousarray(cams, np.int32)
        self._ck(self.lib.psnw_rebuild_pyramids_slot(self._h, len(cams), cams.ctypes.data, slot))

    def lk_stats(self, enable):
        """-> dict(solves, template_passes, iterations) gathered since the previous call; switches the counting kernel on / off."""
        out = (C.c_uint64 * 3)()
        self._ck(self.lib.psnw_lk_stats(self._h, 1 if enable else 0, out))
        return {"solves": int(out[0]), "template_passes": int(out[1]), "iterations": int(out[2])}

    def read_level(self, cam, slot, level):
        w, h, _ = self.level_info(level)
        img = np.empty((h, w), np.uint8)
        dxy = np.empty((h, w, 2), np.int16)
        self._ck(self.lib.psnw_read_level(self._h, cam, slot, level, img.ctypes.data, dxy.ctypes.data))
        return img, dxy

    def read_level_apron(self, cam, slot, level):
        """Image of one level including the REFLECT_101 apron kept in HBM (levels >= 1): (array, apron width)."""
        a = C.c_int(0)
        self._ck(self.lib.psnw_read_level_apron(self._h, cam, slot, level, None, C.byref(a)))
        w, h, _ = self.level_info(level)
        out = np.empty((h + 2 * a.value, w + 2 * a.value), np.uint8)
        self._ck(self.lib.psnw_read_level_apron(self._h, cam, slot, level, out.ctypes.data, C.byref(a)))
        return out, a.value

    def read_min_eig(self, cam=0, slot=0):
        out = np.empty((self.height, self.width), np.float32)
        self._ck(self.lib.psnw_read_min_eig(self._h, cam, slot, out.ctypes.data))
        return out

    # -- the three hot-path operators ------------------------------------------------------------
    def detect(self, boxes, max_per_target=100, quality=0.01, min_dist=3.0, cam=0, slot=0):
        """-> (pts float32 [n_targets, max_per_target, 2], counts int32 [n_targets])."""
        boxes = np.ascontiguousarray(boxes, np.float64).reshape(-1, 4)
        n = boxes.shape[0]
        pts = np.zeros((n, max_per_target, 2), np.float32)
        counts = np.zeros(n, np.int32)
        self._ck(self.lib.psnw_klt_detect(self._h, cam, slot, boxes.ctypes.data, n, max_per_target, quality, min_dist,
                                          pts.ctypes.data, counts.ctypes.data, MEM_HOST))
        return pts, counts

    def gridfast_detect(self, boxes, max_per_target=100, threshold=10, grid=(4, 4), max_total=1000, cam=0, slot=0):
        """GridFAST keypoints (the reference's detector) -> (pts float32 [n_targets, max_per_target, 2], counts)."""
        boxes = np.ascontiguousarray(boxes, np.float64).reshape(-1, 4)
        n = boxes.shape[0]
        pts = np.zeros((n, max_per_target, 2), np.float32)
        counts = np.zeros(n, np.int32)
        self._ck(self.lib.psnw_gridfast_detect(self._h, cam, slot, boxes.ctypes.data, n, max_per_target, threshold, grid[0], grid[1],
                                               max_total, pts.ctypes.data, counts.ctypes.data, MEM_HOST))
        return pts, counts

    def read_fast_map(self, cam=0):
        out = np.empty((self.height, self.width), np.uint8)
        self._ck(self.lib.psnw_read_fast_map(self._h, cam, out.ctypes.data))
        return out

    def track(self, pts, cam=0, slot_from=1, slot_to=0, target_of_pt=None, win_wh=None):
        """-> (out [n,2] f32, status [n] u8, err [n] f32, fb_err [n] f32 or None)."""
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
        n = pts.shape[0]
        tgt = None if target_of_pt is None else np.ascontiguousarray(target_of_pt, np.int32)
        win = None if win_wh is None else np.ascontiguousarray(win_wh, np.int32).reshape(-1, 2)
        n_targets = 1 if win is None else win.shape[0]
        out = np.zeros((n, 2), np.float32)
        status = np.zeros(n, np.uint8)
        err = np.zeros(n, np.float32)
        fb = np.zeros(n, np.float32) if self.do_fb else None
        self._ck(self.lib.psnw_klt_track(self._h, cam, slot_from, slot_to, _ptr(pts), _ptr(tgt), n, _ptr(win), n_targets,
                                         _ptr(out), _ptr(status), _ptr(err), _ptr(fb), MEM_HOST))
        return out, status, err, fb

    def track_update(self, pts, offsets, boxes, win_wh=None, cam=0, slot_from=1, slot_to=0, use_status=True):
        """psnw_klt_track + psnw_box_update as one call -> (out, status, err, boxes_out, inlier_mask, inlier_counts)."""
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
        offsets = np.ascontiguousarray(offsets, np.int32)
        boxes = np.ascontiguousarray(boxes, np.float64).reshape(-1, 4)
        win = None if win_wh is None else np.ascontiguousarray(win_wh, np.int32).reshape(-1, 2)
        n, nt = pts.shape[0], boxes.shape[0]
        out, status, err = np.zeros((n, 2), np.float32), np.zeros(n, np.uint8), np.zeros(n, np.float32)
        bout, mask, cnt = np.zeros((nt, 4), np.float64), np.zeros(n, np.uint8), np.zeros(nt, np.int32)
        self._ck(self.lib.psnw_klt_track_update(self._h, cam, slot_from, slot_to, _ptr(pts), _ptr(offsets), nt, _ptr(win), _ptr(boxes),
                                                1 if use_status else 0, _ptr(out), _ptr(status), _ptr(err), _ptr(bout), _ptr(mask),
                                                _ptr(cnt), MEM_HOST))
        return out, status, err, bout, mask, cnt

    def track_batch(self, jobs, mem=MEM_HOST):
        """jobs: list of dicts with cam, slot_from, slot_to, pts, [target_of_pt], [win_wh], out, status, err, [fb_err]
        (numpy arrays or raw addresses; outputs must be preallocated)."""
        arr = (TrackJob * len(jobs))()
        for a, j in zip(arr, jobs):
            a.cam, a.slot_from, a.slot_to = j.get("cam", 0), j.get("slot_from", 1), j.get("slot_to", 0)
            a.n_pts = j["n_pts"]
            a.pts_xy, a.target_of_pt = _ptr(j["pts"]), _ptr(j.get("target_of_pt"))
            a.n_targets, a.win_wh = j.get("n_targets", 1), _ptr(j.get("win_wh"))
            a.out_xy, a.status, a.err, a.fb_err = _ptr(j["out"]), _ptr(j["status"]), _ptr(j["err"]), _ptr(j.get("fb_err"))
        self._ck(self.lib.psnw_klt_track_batch(self._h, arr, len(jobs), mem))

    def box_update(self, boxes, pre, cur, offsets, status=None, mode=None):
        """-> (boxes_out f64 [n,4], inlier_mask u8 [n_pts], inlier_counts i32 [n]).  mode: BOX_REFERENCE (default) / BOX_MEDIANFLOW"""
        boxes = np.ascontiguousarray(boxes, np.float64).reshape(-1, 4)
        pre = np.ascontiguousarray(pre, np.float32).reshape(-1, 2)
        cur = np.ascontiguousarray(cur, np.float32).reshape(-1, 2)
        offsets = np.ascontiguousarray(offsets, np.int32)
        st = None if status is None else np.ascontiguousarray(status, np.uint8)
        n = boxes.shape[0]
        out = np.zeros((n, 4), np.float64)
        mask = np.zeros(max(pre.shape[0], 1), np.uint8)
        counts = np.zeros(max(n, 1), np.int32)
        self._ck(self.lib.psnw_box_update(self._h, _ptr(boxes), n, _ptr(pre), _ptr(cur), _ptr(st), _ptr(offsets),
                                          BOX_REFERENCE if mode is None else mode, _ptr(out), _ptr(mask), _ptr(counts), MEM_HOST))
        return out, mask[:pre.shape[0]], counts[:n]

    def points_in_boxes(self, pts, tracker_offsets, det_boxes):
        """-> int32 [n_detections, n_trackers]: tracked points of each tracker inside each detection box."""
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
        off = np.ascontiguousarray(tracker_offsets, np.int32)
        boxes = np.ascontiguousarray(det_boxes, np.float64).reshape(-1, 4)
        nt, nd = len(off) - 1, len(boxes)
        counts = np.zeros((max(nd, 1), max(nt, 1)), np.int32)
        self._ck(self.lib.psnw_points_in_boxes(self._h, _ptr(pts), _ptr(off), nt, _ptr(boxes), nd, _ptr(counts), MEM_HOST))
        return counts[:nd, :nt]

    def backward_chain(self, pts, counts, boxes, win_wh=None, n_steps=3, min_inliers=4, cam=0):
        """pts [n_targets, cap, 2] padded per target -> (boxes_steps [n_steps, n, 4], steps_done [n], pts_steps [n_steps, n, cap, 2],
        counts_steps [n_steps, n], first_inlier_mask [n, cap])."""
        pts = np.ascontiguousarray(pts, np.float32)
        n, cap = pts.shape[0], pts.shape[1]
        counts = np.ascontiguousarray(counts, np.int32)
        boxes = np.ascontiguousarray(boxes, np.float64).reshape(-1, 4)
        win = None if win_wh is None else np.ascontiguousarray(win_wh, np.int32).reshape(-1, 2)
        bs, sd = np.zeros((n_steps, n, 4), np.float64), np.zeros(n, np.int32)
        ps, cs = np.zeros((n_steps, n, cap, 2), np.float32), np.zeros((n_steps, n), np.int32)
        fm = np.zeros((n, cap), np.uint8)
        self._ck(self.lib.psnw_klt_backward_chain(self._h, cam, n_steps, _ptr(pts), _ptr(counts), n, cap, _ptr(win), _ptr(boxes), min_inliers,
                                                  _ptr(bs), _ptr(sd), _ptr(ps), _ptr(cs), _ptr(fm), MEM_HOST))
        return bs, sd, ps, cs, fm

    def flow_majority(self, pts, tracker_offsets, tracker_boxes, tracker_ok, det_boxes, ratio=0.5):
        """-> (counts int32 [n_detections, n_trackers], wipe uint8 [n_detections])."""
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
        off = np.ascontiguousarray(tracker_offsets, np.int32)
        tb = np.ascontiguousarray(tracker_boxes, np.float64).reshape(-1, 4)
        ok = np.ascontiguousarray(tracker_ok, np.uint8)
        db = np.ascontiguousarray(det_boxes, np.float64).reshape(-1, 4)
        nt, nd = len(off) - 1, len(db)
        counts, wipe = np.zeros((max(nd, 1), max(nt, 1)), np.int32), np.zeros(max(nd, 1), np.uint8)
        self._ck(self.lib.psnw_flow_majority(self._h, _ptr(pts), _ptr(off), nt, _ptr(tb), _ptr(ok), _ptr(db), nd, ratio, _ptr(counts), _ptr(wipe), MEM_HOST))
        return counts[:nd, :nt], wipe[:nd]

    def box_update_raw(self, boxes_in, n_targets, pre, cur, status, offsets, boxes_out, mask, counts, mem):
        self._ck(self.lib.psnw_box_update(self._h, _ptr(boxes_in), n_targets, _ptr(pre), _ptr(cur), _ptr(status),
                                          _ptr(offsets), BOX_REFERENCE, _ptr(boxes_out), _ptr(mask), _ptr(counts), mem))
