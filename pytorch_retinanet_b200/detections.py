"""Drop-in ``process_detections`` (reference: retinanet/models.py:160-243), whole batch on CUDA.

``process_detections(self, outputs, anchors, im_szs)`` keeps the reference signature: it reads
``self.score_thres / self.nms_thres / self.detections_per_img``, pops ``cls_preds`` and
``bbox_preds`` from ``outputs`` (models.py:168-169) and returns ``List[Dict]`` with ``boxes``
[K,4] fp32, ``scores`` [K] fp32, ``labels`` [K] int64 (1-based), K <= detections_per_img.
Exactly one device->host copy (the per-image detection counts + candidate-pool status) is made.
Final tie rule (the reference's sort is unstable): score desc, class asc, anchor asc.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _native
from .config import BBOX_REG_WEIGHTS, MAX_DETECTIONS_PER_IMAGE, NMS_THRES, SCORE_THRES
from .box_utils import _REG_WEIGHTS_C
from .losses import _shared_anchors

_HW_CACHE: Dict[Tuple, Tensor] = {}


def _image_sizes_tensor(im_szs, dev) -> Tensor:
    """[N,2] int32 (h,w) on the device; cached per (sizes, device) — batches of equal-size images repeat."""
    key = (tuple((int(h), int(w)) for h, w in im_szs), dev.index)
    t = _HW_CACHE.get(key)
    if t is None:
        if len(_HW_CACHE) > 64:
            _HW_CACHE.clear()
        t = torch.tensor(key[0], dtype=torch.int32).reshape(-1, 2).to(dev)
        _HW_CACHE[key] = t
    return t


def default_candidate_capacity(N: int, A: int, C: int) -> int:
    return int(min(N * A * C, max(1 << 20, N * (1 << 16))))


def postprocess_batch(cls_preds: Tensor, bbox_preds: Tensor, anchors: Tensor, anchor_stride: int,
                      im_szs: Sequence[Tuple[int, int]], score_thres: float, nms_thres: float, max_det: int,
                      pre_nms_topk: Optional[int] = None, level_offsets: Optional[Sequence[int]] = None,
                      cand_capacity: Optional[int] = None, algo: str = "auto"):
    """Returns (boxes [N,max_det,4], scores [N,max_det], labels [N,max_det] int64, counts list[int]).

    ``algo``: "auto" = lazy per-image algorithm, transparently repeated with the general
    per-(image,class) algorithm when the lazy one reports it could not finish an image;
    "lazy" / "general" force one of them (tests).  Results are identical."""
    lib = _native.load()
    dev = cls_preds.device
    N, A, C = cls_preds.shape
    x = cls_preds.detach()
    b = bbox_preds.detach()
    x = x if (x.dtype == torch.float32 and x.is_contiguous()) else x.to(torch.float32).contiguous()
    b = b if (b.dtype == torch.float32 and b.is_contiguous()) else b.to(torch.float32).contiguous()
    if len(im_szs) != N:
        raise ValueError(f"{len(im_szs)} image sizes for {N} images")
    hw = _image_sizes_tensor(im_szs, dev)
    out_boxes = torch.empty((N, max_det, 4), dtype=torch.float32, device=dev)
    out_scores = torch.empty((N, max_det), dtype=torch.float32, device=dev)
    out_labels = torch.empty((N, max_det), dtype=torch.int64, device=dev)
    meta = torch.empty((N + 4,), dtype=torch.int32, device=dev)   # counts [N] + status [4]
    use_general = algo == "general" or (A * C >= (1 << 32))
    cap = int(cand_capacity) if cand_capacity else default_candidate_capacity(N, A, C)
    topk = int(pre_nms_topk) if pre_nms_topk else 0
    lvl = None
    nlev = 0
    if topk:
        if level_offsets is None:
            raise ValueError("pre_nms_topk requires level_offsets")
        nlev = len(level_offsets) - 1
        lvl = (ctypes.c_int64 * len(level_offsets))(*[int(v) for v in level_offsets])
    while True:
        ws_bytes = lib.rn_postprocess_workspace_bytes(N, A, C, cap, max_det)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = lib.rn_postprocess(_native.ptr(x, torch.float32, "cls_preds"), _native.ptr(b, torch.float32, "bbox_preds"),
                                    _native.ptr(anchors, torch.float32, "anchors"), anchor_stride, _native.ptr(hw), N, A, C,
                                    float(score_thres), float(nms_thres), int(max_det),
                                    _REG_WEIGHTS_C, topk, lvl, nlev, 1 if use_general else 0, cap,
                                    _native.ptr(out_boxes), _native.ptr(out_scores), _native.ptr(out_labels),
                                    meta.data_ptr(), meta.data_ptr() + 4 * N, _native.ptr(ws), ws_bytes,
                                    _native.stream_ptr(dev))
        _native.check(rc, "rn_postprocess")
        host = meta.tolist()          # the single D2H copy / sync of the path
        found, capacity, fallback = host[N], host[N + 1], host[N + 2]
        if found > capacity:
            cap = found               # candidate pool overflowed: the exact need is now known
            continue
        if fallback and not use_general:
            if algo == "lazy":
                raise _native.NativeError("rn_postprocess: lazy algorithm could not finish (algo='lazy' forced)")
            use_general = True        # rare: an image needs more rounds than the lazy budget
            continue
        return out_boxes, out_scores, out_labels, host[:N]


def process_detections(self, outputs: Dict[str, Tensor], anchors: List[Tensor],
                       im_szs: List[Tuple[int, int]]) -> List[Dict[str, Tensor]]:
    class_logits = outputs.pop("cls_preds")
    bboxes = outputs.pop("bbox_preds")
    an, stride = _shared_anchors(anchors)
    score_thres = getattr(self, "score_thres", SCORE_THRES)
    nms_thres = getattr(self, "nms_thres", NMS_THRES)
    max_det = getattr(self, "detections_per_img", MAX_DETECTIONS_PER_IMAGE)
    topk = getattr(self, "pre_nms_topk", None)
    lvl = getattr(self, "anchor_level_offsets", None)
    ob, os_, ol, counts = postprocess_batch(class_logits, bboxes, an, stride, im_szs, score_thres, nms_thres,
                                            max_det, topk, lvl)
    return [{"boxes": ob[i, :k], "scores": os_[i, :k], "labels": ol[i, :k]} for i, k in enumerate(counts)]
