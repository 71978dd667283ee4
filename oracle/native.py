"""ctypes access to oracle/c/oracle_kernels.c (test infrastructure; see oracle/__init__.py)."""
import ctypes

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
        _lib.oracle_nms.restype = ctypes.c_int64
        _lib.oracle_nms.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                    ctypes.c_double, ctypes.c_void_p]
        _lib.oracle_roi_align.restype = None
        _lib.oracle_roi_align.argtypes = [ctypes.c_void_p] + [ctypes.c_int64] * 4 + [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_float, ctypes.c_int, ctypes.c_int,
            ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        _lib.oracle_argsort_desc_stable.restype = None
        _lib.oracle_argsort_desc_stable.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def nms(boxes, scores, iou_threshold: float) -> np.ndarray:
    """tv:ops/boxes.py:20-48 -> torchvision::nms CPU (tv-csrc:ops/cpu/nms_kernel.cpp:116)."""
    boxes = _f32(boxes).reshape(-1, 4)
    scores = _f32(scores).reshape(-1)
    n = boxes.shape[0]
    keep = np.empty(max(n, 1), dtype=np.int64)
    cnt = lib().oracle_nms(boxes.ctypes.data, scores.ctypes.data, n, float(iou_threshold),
                           keep.ctypes.data)
    return keep[:cnt].copy()


def argsort_desc_stable(scores) -> np.ndarray:
    scores = _f32(scores).reshape(-1)
    order = np.empty(max(scores.shape[0], 1), dtype=np.int64)
    lib().oracle_argsort_desc_stable(scores.ctypes.data, scores.shape[0], order.ctypes.data)
    return order[:scores.shape[0]].copy()


def roi_align(inp, rois, spatial_scale: float, pooled_h: int, pooled_w: int,
              sampling_ratio: int, aligned: bool) -> np.ndarray:
    """tv:ops/roi_align.py:203-260 -> torchvision::roi_align CPU
    (tv-csrc:ops/cpu/roi_align_kernel.cpp:393). rois [K,5] = (batch, x1, y1, x2, y2)."""
    inp = _f32(inp)
    rois = _f32(rois).reshape(-1, 5)
    n, c, h, w = inp.shape
    k = rois.shape[0]
    out = np.zeros((k, c, pooled_h, pooled_w), dtype=np.float32)
    if k:
        lib().oracle_roi_align(inp.ctypes.data, n, c, h, w, rois.ctypes.data, k,
                               float(spatial_scale), pooled_h, pooled_w, int(sampling_ratio),
                               int(bool(aligned)), out.ctypes.data)
    return out
