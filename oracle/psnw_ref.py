"""ctypes front-end of oracle/_ref/libpsnw_ref.so: the reference's OWN LocalSearchKLT and PSN_Rect helpers, compiled
unchanged from /root/reference/psn_where/PSNWhere_Tracker2D.cpp:452-554 and PSNWhere_Types.h:18-182 behind a stub of
the OpenCV types they touch (oracle/ref_shim/, recipe in oracle/Makefile).  TEST INFRASTRUCTURE ONLY -- the pin of the
oracle's and the CUDA kernel's box motion to the reference's own code.  Only tests/ may import this module.

The library is built where the reference tree exists (this container); the prebuilt .so travels to the GPU box
(oracle/_ref/ is git-ignored, not gpurun-ignored), where /root/reference does not exist and nothing reads it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libpsnw_ref.so")
REF_SRC = "/root/reference/psn_where/PSNWhere_Tracker2D.cpp"

_lib = None


def available():
    return os.path.exists(LIB_PATH) or os.path.exists(REF_SRC)


def lib():
    global _lib
    if _lib is None:
        if os.path.exists(REF_SRC):          # (re)build from the reference sources where they lie
            subprocess.check_call(["make", "-C", _HERE, "_ref/libpsnw_ref.so"], stdout=subprocess.DEVNULL)
        if not os.path.exists(LIB_PATH):
            raise ImportError("oracle/_ref/libpsnw_ref.so is missing and /root/reference is not present to build it")
        _lib = C.CDLL(LIB_PATH)
        _lib.psnref_local_search_klt.restype = C.c_int
        _lib.psnref_overlapped_area.restype = C.c_double
        _lib.psnref_contain.restype = C.c_int
        _lib.psnref_contain.argtypes = [C.c_void_p, C.c_float, C.c_float]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def local_search_klt(box, pre, cur):
    """CPSNWhere_Tracker2D::LocalSearchKLT(box, pre, cur, inlierIdx) -> (PSN_Rect as float64[4], inlier indices)."""
    box = np.ascontiguousarray(box, np.float64).reshape(4)
    pre = np.ascontiguousarray(pre, np.float32).reshape(-1, 2)
    cur = np.ascontiguousarray(cur, np.float32).reshape(-1, 2)
    n = pre.shape[0]
    out = np.zeros(4, np.float64)
    idx = np.zeros(max(n, 1), np.int32)
    k = lib().psnref_local_search_klt(_p(box), _p(pre), _p(cur), C.c_int(n), _p(out), _p(idx))
    return out, idx[:k].copy()


def overlapped_area(a, b):
    a = np.ascontiguousarray(a, np.float64).reshape(4)
    b = np.ascontiguousarray(b, np.float64).reshape(4)
    return float(lib().psnref_overlapped_area(_p(a), _p(b)))


def contain(box, x, y):
    box = np.ascontiguousarray(box, np.float64).reshape(4)
    return bool(lib().psnref_contain(_p(box), C.c_float(x), C.c_float(y)))


def crop_roi(box, width, height):
    """PSN_Rect::cropWithSize(width, height).cv() -> (x, y, w, h) ints (the mask ROI of Tracker2D.cpp:735)."""
    box = np.ascontiguousarray(box, np.float64).reshape(4)
    roi = np.zeros(4, np.int32)
    lib().psnref_crop_roi(_p(box), C.c_double(width), C.c_double(height), _p(roi))
    return tuple(int(v) for v in roi)
