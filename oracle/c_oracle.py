"""TEST INFRASTRUCTURE ONLY — ctypes binding of ``oracle/oracle.c`` (see its header).

Allowed callers: ``tests/``, ``__graft_entry__.smoke()``, ``bench.py`` (cpu_baseline legs).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib: Optional[ctypes.CDLL] = None

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i64p = ctypes.POINTER(ctypes.c_int64)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_level_anchors.restype = ctypes.c_int64
        _lib.orc_nms.restype = ctypes.c_int64
        _lib.orc_postprocess_image.restype = ctypes.c_int64
        _lib.orc_iou.restype = ctypes.c_float
    return _lib


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a, t):
    return a.ctypes.data_as(t)


def cell_anchors(sizes, ratios) -> np.ndarray:
    s, r = np.asarray(sizes, dtype=np.float64), np.asarray(ratios, dtype=np.float64)
    out = np.empty((len(s) * len(r), 4), dtype=np.float32)
    lib().orc_cell_anchors(_p(s, _f64p), len(s), _p(r, _f64p), len(r), _p(out, _f32p))
    return out


def image_anchors(grid_sizes, strides, sizes, ratios, offset=0.0) -> np.ndarray:
    parts = []
    for (H, W), st, sz, rt in zip(grid_sizes, strides, sizes, ratios):
        cell = cell_anchors(sz, rt)
        out = np.empty((H * W * cell.shape[0], 4), dtype=np.float32)
        lib().orc_level_anchors(int(H), int(W), int(st), ctypes.c_double(offset), _p(cell, _f32p),
                                cell.shape[0], _p(out, _f32p))
        parts.append(out)
    return np.concatenate(parts)


def match(anchors, gt, fg_thr=0.5, bg_thr=0.4) -> np.ndarray:
    a, g = _f32(anchors).reshape(-1, 4), _f32(gt).reshape(-1, 4)
    out = np.empty(a.shape[0], dtype=np.int64)
    lib().orc_match(_p(a, _f32p), ctypes.c_int64(a.shape[0]), _p(g, _f32p), ctypes.c_int64(g.shape[0]),
                    ctypes.c_float(fg_thr), ctypes.c_float(bg_thr), _p(out, _i64p))
    return out


def encode(gt, anchors, weights=(1.0, 1.0, 1.0, 1.0)) -> np.ndarray:
    g, a, w = _f32(gt).reshape(-1, 4), _f32(anchors).reshape(-1, 4), _f32(weights)
    out = np.empty_like(g)
    lib().orc_encode(_p(g, _f32p), _p(a, _f32p), _p(w, _f32p), ctypes.c_int64(g.shape[0]), _p(out, _f32p))
    return out


def decode(act, anchors, weights=(1.0, 1.0, 1.0, 1.0)) -> np.ndarray:
    d, a, w = _f32(act).reshape(-1, 4), _f32(anchors).reshape(-1, 4), _f32(weights)
    out = np.empty_like(d)
    lib().orc_decode(_p(d, _f32p), _p(a, _f32p), _p(w, _f32p), ctypes.c_int64(d.shape[0]), _p(out, _f32p))
    return out


def image_loss(anchors, logits, bbox, gt, labels, fg_thr=0.5, bg_thr=0.4, alpha=0.25, gamma=2.0, beta=0.1,
               weights=(1.0, 1.0, 1.0, 1.0)) -> Tuple[float, float, int, np.ndarray]:
    """Returns (cls_loss_i, reg_loss_i, F_i, matches) for one image."""
    a, x, b = _f32(anchors).reshape(-1, 4), _f32(logits), _f32(bbox).reshape(-1, 4)
    g = _f32(gt).reshape(-1, 4)
    lab = np.ascontiguousarray(np.asarray(labels, dtype=np.int64))
    w = _f32(weights)
    m = np.empty(a.shape[0], dtype=np.int64)
    out3 = np.zeros(3, dtype=np.float64)
    lib().orc_image_loss(_p(a, _f32p), ctypes.c_int64(a.shape[0]), _p(x, _f32p), int(x.shape[-1]), _p(b, _f32p),
                         _p(g, _f32p), _p(lab, _i64p), ctypes.c_int64(g.shape[0]), ctypes.c_float(fg_thr),
                         ctypes.c_float(bg_thr), ctypes.c_float(alpha), ctypes.c_float(gamma), ctypes.c_float(beta),
                         _p(w, _f32p), _p(m, _i64p), _p(out3, _f64p))
    return float(out3[0]), float(out3[1]), int(out3[2]), m


def nms(boxes, scores, thr: float) -> np.ndarray:
    b, s = _f32(boxes).reshape(-1, 4), _f32(scores).reshape(-1)
    keep = np.empty(b.shape[0], dtype=np.int64)
    n = lib().orc_nms(_p(b, _f32p), _p(s, _f32p), ctypes.c_int64(b.shape[0]), ctypes.c_double(thr), _p(keep, _i64p))
    return keep[:n].copy()


def postprocess_image(logits, bbox, anchors, im_hw, score_thr=0.05, nms_thr=0.5, max_det=100,
                      weights=(1.0, 1.0, 1.0, 1.0)):
    x, b, a, w = _f32(logits), _f32(bbox).reshape(-1, 4), _f32(anchors).reshape(-1, 4), _f32(weights)
    ob = np.empty((max_det, 4), dtype=np.float32)
    os_ = np.empty(max_det, dtype=np.float32)
    ol = np.empty(max_det, dtype=np.int64)
    n = lib().orc_postprocess_image(_p(x, _f32p), _p(b, _f32p), _p(a, _f32p), ctypes.c_int64(a.shape[0]),
                                    int(x.shape[-1]), int(im_hw[0]), int(im_hw[1]), ctypes.c_float(score_thr),
                                    ctypes.c_double(nms_thr), ctypes.c_int64(max_det), _p(w, _f32p),
                                    _p(ob, _f32p), _p(os_, _f32p), _p(ol, _i64p))
    return ob[:n].copy(), os_[:n].copy(), ol[:n].copy()
