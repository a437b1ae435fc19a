"""ctypes front-end of the C++ host mirror of CPSNWhere_Tracker2D (integration/psnw_tracker2d.*).

The class itself is C++ (the reference is compiled code); this module only lets the Python tests and examples drive
it: Tracker2D(ctx, cam).run(boxes, gray, frame_idx) == CPSNWhere_Tracker2D::Run (psn_where/PSNWhere_Tracker2D.cpp:251).
"""
import ctypes as C
import os

import numpy as np

from mcmtt_opticalflow_b200 import klt

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libpsnw_tracker2d.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        klt.load()
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: build it with `make -C integration`" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        lib.psnt_create.restype = C.c_void_p
        lib.psnt_create.argtypes = [C.c_void_p, C.c_uint, C.c_int, C.c_int, C.c_int]
        lib.psnt_destroy.argtypes = [C.c_void_p]
        lib.psnt_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_uint]
        lib.psnt_last_error.restype = C.c_char_p
        lib.psnt_last_error.argtypes = [C.c_void_p]
        lib.psnt_get_objects.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.psnt_get_detection_boxes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.psnt_hungarian.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        lib.psnt_read_detections.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.psnt_write_result.argtypes = [C.c_void_p, C.c_char_p]
        lib.psnt_result_roundtrip.argtypes = [C.c_char_p, C.c_uint, C.c_uint, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                              C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = lib
    return _lib


def hungarian(cost):
    """Minimum-cost assignment on a float32 [rows, cols] matrix -> (rows, cols) index arrays."""
    cost = np.ascontiguousarray(cost, np.float32)
    r, c = cost.shape
    mr, mc = np.zeros(max(r, c), np.int32), np.zeros(max(r, c), np.int32)
    n = load().psnt_hungarian(cost.ctypes.data, r, c, mr.ctypes.data, mc.ctypes.data)
    return mr[:n].copy(), mc[:n].copy()


def read_detections(path, cap=256):
    """PETS.S2.L1 full-body detection file (psn::ReadDetectionResultWithTxt, psn_where/PSNWhere_Utils.cpp:996-1088)
    -> (boxes float64 [n,4], heads float64 [n,4]); a missing file gives empty arrays."""
    boxes, heads = np.zeros((cap, 4), np.float64), np.zeros((cap, 4), np.float64)
    n = load().psnt_read_detections(os.fsencode(path), boxes.ctypes.data, heads.ctypes.data, cap)
    return boxes[:n].copy(), heads[:n].copy()


class Tracker2D:
    """One CPSNWhere_Tracker2D instance (one camera) on a shared klt.KLTContext."""

    def __init__(self, ctx, cam_index=0, cam_id=0):
        self.lib, self.ctx = load(), ctx
        self._h = self.lib.psnt_create(ctx._h, cam_id, cam_index, ctx.width, ctx.height)

    def close(self):
        if self._h:
            self.lib.psnt_destroy(self._h)
            self._h = None

    def run(self, boxes, gray, frame_idx, heads=None):
        """-> (ids uint32 [k], boxes float64 [k,4], n_features int32 [k]) of stTrack2DResult.object2DInfos."""
        boxes = np.ascontiguousarray(boxes, np.float64).reshape(-1, 4)
        heads = None if heads is None else np.ascontiguousarray(heads, np.float64).reshape(-1, 4)
        gray = np.ascontiguousarray(gray, np.uint8)
        n = self.lib.psnt_run(self._h, boxes.ctypes.data, None if heads is None else heads.ctypes.data, len(boxes),
                              gray.ctypes.data, gray.strides[0], frame_idx)
        if n < 0:
            raise RuntimeError((self.lib.psnt_last_error(self._h) or b"").decode())
        ids, ob, nf = np.zeros(max(n, 1), np.uint32), np.zeros((max(n, 1), 4), np.float64), np.zeros(max(n, 1), np.int32)
        self.lib.psnt_get_objects(self._h, ids.ctypes.data, ob.ctypes.data, nf.ctypes.data, n)
        return ids[:n], ob[:n], nf[:n]

    def write_result(self, path):
        """track2D_result_cam%d_frame%04d.txt of the last run (CPSNWhere_Tracker2D::FilePrintResult, Tracker2D.cpp:1268-1342)."""
        if self.lib.psnt_write_result(self._h, os.fsencode(path)) != 0:
            raise OSError("cannot write %s" % path)

    def detection_boxes(self, cap=512):
        """Backward-chain boxes (current -> past) of the detections kept by the last run."""
        nb, bx = np.zeros(cap, np.int32), np.zeros((cap, 4, 4), np.float64)
        n = self.lib.psnt_get_detection_boxes(self._h, nb.ctypes.data, bx.ctypes.data, cap)
        return [bx[i, :nb[i]].copy() for i in range(n)]
