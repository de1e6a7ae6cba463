"""ORACLE — test infrastructure only.  ctypes front end of the plain-C restatement in `oracle/c/`
(`roi_align_ref.c`, `nms_ref.c`); numpy in, numpy out.  Built by `oracle/c/Makefile` into
`oracle/_build/liboracle.so` (git-ignored; travels to the GPU box with the snapshot)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, "c", f) for f in ("roi_align_ref.c", "nms_ref.c", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-s", "-B", "-C", os.path.join(_HERE, "c")])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        fp, ip, c_int, c_float = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        for name in ("oracle_roi_align_fwd", "oracle_roi_align_bwd"):
            f = getattr(L, name)
            f.argtypes = [fp, fp, fp] + [c_int] * 7 + [c_float, c_int, c_int]
            f.restype = c_int
        L.oracle_batched_nms.argtypes = [fp, fp, ip, ctypes.c_int64, ctypes.c_double, c_int, ip]
        L.oracle_batched_nms.restype = ctypes.c_int64
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def roi_align_fwd(inp, rois, output_size, spatial_scale, sampling_ratio, aligned=True):
    inp, rois = _f32(inp), _f32(rois).reshape(-1, 5)
    n, c, h, w = inp.shape
    ph, pw = output_size
    out = np.empty((rois.shape[0], c, ph, pw), dtype=np.float32)
    lib().oracle_roi_align_fwd(inp.ctypes.data, rois.ctypes.data, out.ctypes.data, n, c, h, w, rois.shape[0],
                               ph, pw, float(spatial_scale), int(sampling_ratio), int(bool(aligned)))
    return out


def roi_align_bwd(gout, rois, in_shape, spatial_scale, sampling_ratio, aligned=True):
    gout, rois = _f32(gout), _f32(rois).reshape(-1, 5)
    n, c, h, w = in_shape
    r, c2, ph, pw = gout.shape
    assert c2 == c and r == rois.shape[0]
    gin = np.empty((n, c, h, w), dtype=np.float32)
    lib().oracle_roi_align_bwd(gout.ctypes.data, rois.ctypes.data, gin.ctypes.data, n, c, h, w, r, ph, pw,
                               float(spatial_scale), int(sampling_ratio), int(bool(aligned)))
    return gin


def batched_nms(boxes, scores, idxs, iou_threshold, coord_trick=None):
    """`coord_trick=None` follows torchvision 0.26's CPU rule (boxes.py:80-83): numel > 4000 -> per-class
    loop, else coordinate offsets."""
    boxes, scores = _f32(boxes).reshape(-1, 4), _f32(scores)
    m = boxes.shape[0]
    keep = np.empty((max(m, 1),), dtype=np.int64)
    if idxs is None:
        ip, trick = None, 0
    else:
        idxs = np.ascontiguousarray(idxs, dtype=np.int64)
        ip = idxs.ctypes.data
        trick = int(boxes.size <= 4000) if coord_trick is None else int(bool(coord_trick))
    nk = lib().oracle_batched_nms(boxes.ctypes.data, scores.ctypes.data, ip, m, float(iou_threshold), trick,
                                  keep.ctypes.data)
    return keep[:nk].copy()
