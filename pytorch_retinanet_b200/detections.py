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
from .config import MAX_DETECTIONS_PER_IMAGE, NMS_THRES, SCORE_THRES
from .box_utils import _REG_WEIGHTS_C
from .losses import _f32_contig, _level_desc, _ptr_array, _shared_anchors

_HW_CACHE: Dict[Tuple, Tensor] = {}
_WS_CACHE: Dict[Tuple, Tensor] = {}


def _workspace(nbytes: int, dev) -> Tensor:
    """Scratch buffer reused across calls on the same (device, stream): the pipeline is stream-ordered,
    so the next call may overwrite it.  Grows monotonically."""
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    t = _WS_CACHE.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        _WS_CACHE[key] = t
    return t


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


_RATIO_CACHE: Dict[Tuple, Tensor] = {}


def _resize_ratio_tensor(im_szs, original_image_sizes, dev) -> Optional[Tensor]:
    """[N,2] fp32 (ratio_h, ratio_w) = fp32(original) / fp32(resized), exactly torchvision's resize_boxes ratios
    (tv:models/detection/transform.py:307-311); None when no resize is requested."""
    if original_image_sizes is None:
        return None
    if len(original_image_sizes) != len(im_szs):
        raise ValueError("original_image_sizes and im_szs differ in length")
    key = (tuple((int(h), int(w)) for h, w in im_szs), tuple((int(h), int(w)) for h, w in original_image_sizes), dev.index)
    t = _RATIO_CACHE.get(key)
    if t is None:
        if len(_RATIO_CACHE) > 64:
            _RATIO_CACHE.clear()
        new = torch.tensor(key[1], dtype=torch.float32).reshape(-1, 2)
        old = torch.tensor(key[0], dtype=torch.float32).reshape(-1, 2)
        t = (new / old).to(dev)
        _RATIO_CACHE[key] = t
    return t


_FORMATS = {"xyxy": 0, "xywh": 1}


def default_candidate_capacity(N: int, A: int, C: int) -> int:
    return int(min(N * A * C, max(1 << 20, N * (1 << 16))))


class PendingDetections:
    """Handle of an enqueued post-processing call.  All kernels and the (pinned, asynchronous) copy of
    the detection counts are already on the stream; ``result()`` waits for that copy only, handles the
    rare re-runs (candidate pool overflow / lazy -> general fallback) and slices the ragged outputs.
    Lets a caller enqueue the next batch before looking at this one (no GPU idle gap at the sync)."""

    def __init__(self, args: dict):
        self._a = args
        self._done = None
        self._launch()

    def _launch(self):
        a = self._a
        lib = _native.load()
        dev, N, A, C = a["dev"], a["N"], a["A"], a["C"]
        meta = torch.empty((N + 4,), dtype=torch.int32, device=dev)   # counts [N] + status [4]
        outs = (_native.ptr(a["out_boxes"]), _native.ptr(a["out_scores"]), _native.ptr(a["out_labels"]),
                meta.data_ptr(), meta.data_ptr() + 4 * N)
        algo_id = 1 if a["use_general"] else 0
        with _native.on_device(dev):
            if a.get("levels") is not None:               # raw per-level conv outputs (row N1)
                xs, bs, desc = a["levels"]
                ws_bytes = lib.rn_postprocess_levels_workspace_bytes(N, A, C, a["cap"], a["max_det"])
                ws = _workspace(ws_bytes, dev)
                rc = lib.rn_postprocess_levels(_ptr_array(xs), _ptr_array(bs), desc, len(xs),
                                               _native.ptr(a["anchors"], torch.float32, "anchors"), a["anchor_stride"],
                                               _native.ptr(a["hw"]), N, A, C, a["score_thres"], a["nms_thres"], a["max_det"],
                                               _REG_WEIGHTS_C, a["topk"], algo_id, a["cap"], *outs,
                                               _native.ptr(ws), ws_bytes, _native.stream_ptr(dev),
                                               _native.ptr(a["ratio"]), a["fmt"])
                what = "rn_postprocess_levels"
            else:
                ws_bytes = lib.rn_postprocess_workspace_bytes(N, A, C, a["cap"], a["max_det"])
                ws = _workspace(ws_bytes, dev)
                rc = lib.rn_postprocess(_native.ptr(a["x"], torch.float32, "cls_preds"),
                                        _native.ptr(a["b"], torch.float32, "bbox_preds"),
                                        _native.ptr(a["anchors"], torch.float32, "anchors"), a["anchor_stride"],
                                        _native.ptr(a["hw"]), N, A, C, a["score_thres"], a["nms_thres"], a["max_det"],
                                        _REG_WEIGHTS_C, a["topk"], a["lvl"], a["nlev"], algo_id, a["cap"], *outs,
                                        _native.ptr(ws), ws_bytes, _native.stream_ptr(dev),
                                        _native.ptr(a["ratio"]), a["fmt"])
                what = "rn_postprocess"
        _native.check(rc, what)
        self._host = torch.empty((N + 4,), dtype=torch.int32, pin_memory=True)
        self._host.copy_(meta, non_blocking=True)       # the single D2H copy of the path
        self._event = torch.cuda.Event()
        self._event.record(torch.cuda.current_stream(dev))

    def result(self):
        """(boxes [N,max_det,4], scores [N,max_det], labels [N,max_det] int64, counts list[int])."""
        if self._done is not None:
            return self._done
        a = self._a
        N = a["N"]
        while True:
            self._event.synchronize()
            host = self._host.tolist()
            found, capacity, fallback = host[N], host[N + 1], host[N + 2]
            if found > capacity:
                a["cap"] = found          # candidate pool overflowed: the exact need is now known
            elif fallback and not a["use_general"]:
                if a["algo"] == "lazy":
                    raise _native.NativeError("rn_postprocess: lazy algorithm could not finish (algo='lazy' forced)")
                a["use_general"] = True   # rare: an image needs more rounds than the lazy budget
            else:
                self._done = (a["out_boxes"], a["out_scores"], a["out_labels"], host[:N])
                return self._done
            self._launch()

    def coco_results(self, image_ids: Sequence[int]) -> List[dict]:
        """COCO-format results of the batch (row N4): what ``CocoEvaluator.prepare_for_coco_detection``
        (utils/coco/coco_eval.py:71-93) builds with three ``.tolist()`` syncs per image, from THREE device->host
        copies for the whole batch.  Requires ``box_format="xywh"``."""
        if self._a["fmt"] != 1:
            raise ValueError("coco_results() needs box_format='xywh'")
        ob, os_, ol, counts = self.result()
        hb, hs, hl = ob.cpu().tolist(), os_.cpu().tolist(), ol.cpu().tolist()
        out = []
        for i, (img, k) in enumerate(zip(image_ids, counts)):
            out.extend({"image_id": img, "category_id": hl[i][j], "bbox": hb[i][j], "score": hs[i][j]} for j in range(k))
        return out

    def detections(self) -> List[Dict[str, Tensor]]:
        return slice_detections(*self.result())


def slice_detections(ob: Tensor, os_: Tensor, ol: Tensor, counts: Sequence[int]) -> List[Dict[str, Tensor]]:
    """Padded slabs [N,M,*] + per-image counts -> the reference's ``List[Dict]`` (models.py:236-242) with THREE view
    calls for the whole batch: every image contributes the sizes (k, M-k) to one ``split_with_sizes`` per field."""
    N, M = ob.shape[0], ob.shape[1]
    sizes = []
    for k in counts:
        sizes.append(k)
        sizes.append(M - k)
    b = ob.view(N * M, -1).split_with_sizes(sizes)
    s = os_.view(N * M).split_with_sizes(sizes)
    l = ol.view(N * M).split_with_sizes(sizes)
    return [{"boxes": b[2 * i], "scores": s[2 * i], "labels": l[2 * i]} for i in range(N)]


def postprocess_batch_async(cls_preds: Tensor, bbox_preds: Tensor, anchors: Tensor, anchor_stride: int,
                            im_szs: Sequence[Tuple[int, int]], score_thres: float, nms_thres: float, max_det: int,
                            pre_nms_topk: Optional[int] = None, level_offsets: Optional[Sequence[int]] = None,
                            cand_capacity: Optional[int] = None, algo: str = "auto",
                            original_image_sizes: Optional[Sequence[Tuple[int, int]]] = None,
                            box_format: str = "xyxy") -> PendingDetections:
    """Enqueues the whole post-processing of a batch and returns without synchronising.

    ``algo``: "auto" = lazy per-image algorithm, transparently repeated with the general
    per-(image,class) algorithm when the lazy one reports it could not finish an image;
    "lazy" / "general" force one of them (tests).  Results are identical."""
    dev = cls_preds.device
    N, A, C = cls_preds.shape
    x = cls_preds.detach()
    b = bbox_preds.detach()
    x = x if (x.dtype == torch.float32 and x.is_contiguous()) else x.to(torch.float32).contiguous()
    b = b if (b.dtype == torch.float32 and b.is_contiguous()) else b.to(torch.float32).contiguous()
    if len(im_szs) != N:
        raise ValueError(f"{len(im_szs)} image sizes for {N} images")
    topk = int(pre_nms_topk) if pre_nms_topk else 0
    lvl, nlev = None, 0
    if topk:
        if level_offsets is None:
            raise ValueError("pre_nms_topk requires level_offsets")
        nlev = len(level_offsets) - 1
        lvl = (ctypes.c_int64 * len(level_offsets))(*[int(v) for v in level_offsets])
    if topk and (algo == "general" or A * C >= (1 << 32)):
        raise ValueError("pre_nms_topk is implemented by the lazy algorithm only (needs A*C < 2^32)")
    args = dict(dev=dev, N=N, A=A, C=C, x=x, b=b, anchors=anchors, anchor_stride=anchor_stride,
                ratio=_resize_ratio_tensor(im_szs, original_image_sizes, dev), fmt=_FORMATS[box_format],
                hw=_image_sizes_tensor(im_szs, dev), score_thres=float(score_thres), nms_thres=float(nms_thres),
                max_det=int(max_det), topk=topk, lvl=lvl, nlev=nlev, algo=algo,
                use_general=(algo == "general" or (A * C >= (1 << 32))),
                cap=int(cand_capacity) if cand_capacity else default_candidate_capacity(N, A, C),
                out_boxes=torch.empty((N, max_det, 4), dtype=torch.float32, device=dev),
                out_scores=torch.empty((N, max_det), dtype=torch.float32, device=dev),
                out_labels=torch.empty((N, max_det), dtype=torch.int64, device=dev))
    return PendingDetections(args)


def postprocess_levels_async(cls_levels: Sequence[Tensor], bbox_levels: Sequence[Tensor], num_classes: int, anchors: Tensor,
                             anchor_stride: int, im_szs: Sequence[Tuple[int, int]], score_thres: float, nms_thres: float,
                             max_det: int, pre_nms_topk: Optional[int] = None, cand_capacity: Optional[int] = None,
                             algo: str = "auto", original_image_sizes: Optional[Sequence[Tuple[int, int]]] = None,
                             box_format: str = "xyxy") -> PendingDetections:
    """:func:`postprocess_batch_async` on the RAW per-level conv outputs."""
    xs, bs = [_f32_contig(t) for t in cls_levels], [_f32_contig(t) for t in bbox_levels]
    desc, A, N = _level_desc(xs, bs, num_classes)
    dev, C = xs[0].device, num_classes
    if len(im_szs) != N:
        raise ValueError(f"{len(im_szs)} image sizes for {N} images")
    topk = int(pre_nms_topk) if pre_nms_topk else 0
    if topk and (algo == "general" or A * C >= (1 << 32)):
        raise ValueError("pre_nms_topk is implemented by the lazy algorithm only (needs A*C < 2^32)")
    args = dict(dev=dev, N=N, A=A, C=C, levels=(xs, bs, desc), anchors=anchors, anchor_stride=anchor_stride,
                ratio=_resize_ratio_tensor(im_szs, original_image_sizes, dev), fmt=_FORMATS[box_format],
                hw=_image_sizes_tensor(im_szs, dev), score_thres=float(score_thres), nms_thres=float(nms_thres),
                max_det=int(max_det), topk=topk, lvl=None, nlev=0, algo=algo,
                use_general=(algo == "general" or (A * C >= (1 << 32))),
                cap=int(cand_capacity) if cand_capacity else default_candidate_capacity(N, A, C),
                out_boxes=torch.empty((N, max_det, 4), dtype=torch.float32, device=dev),
                out_scores=torch.empty((N, max_det), dtype=torch.float32, device=dev),
                out_labels=torch.empty((N, max_det), dtype=torch.int64, device=dev))
    return PendingDetections(args)


def postprocess_batch(*args, **kw):
    """Synchronous form of :func:`postprocess_batch_async`: returns
    (boxes [N,max_det,4], scores [N,max_det], labels [N,max_det] int64, counts list[int])."""
    return postprocess_batch_async(*args, **kw).result()


def _level_offsets(self):
    """Anchor offsets of the pyramid levels, needed only when ``self.pre_nms_topk`` is set: an explicit
    ``self.anchor_level_offsets`` wins, else the ones recorded by our AnchorGenerator's last forward."""
    if not getattr(self, "pre_nms_topk", None):
        return None
    offs = getattr(self, "anchor_level_offsets", None)
    if offs is None:
        offs = getattr(getattr(self, "anchor_generator", None), "last_level_offsets", None)
    return offs


class _GraphPending:
    """Result handle of ``process_detections`` in graph mode (``self.rn_graph = True``): same interface as
    :class:`PendingDetections`; the outputs are copied out of the graph's static buffers (44 kB for 16 images)."""

    def __init__(self, step_result, cls_preds, bbox_preds):
        self._r, self._x, self._b = step_result, cls_preds, bbox_preds
        self._done = None

    def result(self):
        if self._done is None:
            ob, os_, ol, counts = self._r.result(self._x, self._b)
            self._done = (ob.clone(), os_.clone(), ol.clone(), counts)
            self._x = self._b = None
        return self._done

    def detections(self) -> List[Dict[str, Tensor]]:
        return slice_detections(*self.result())


def _graph_detections(self, x: Tensor, b: Tensor, an: Tensor, im_szs, original_image_sizes, box_format):
    """Graph replay of the inference half (opt-in, ``patch_retinanet(model, graph=True)``), keyed like
    ``RetinaNetLosses(graph=True)`` on the input addresses and shapes; None when the call does not qualify."""
    if not (x.is_cuda and x.dtype == torch.float32 and b.dtype == torch.float32 and x.is_contiguous() and b.is_contiguous()
            and not getattr(self, "pre_nms_topk", None)):
        return None
    cache = self.__dict__.setdefault("_rn_det_graphs", {})
    key = (x.data_ptr(), b.data_ptr(), tuple(x.shape), an.data_ptr(), tuple((int(h), int(w)) for h, w in im_szs),
           None if original_image_sizes is None else tuple((int(h), int(w)) for h, w in original_image_sizes), box_format,
           float(getattr(self, "score_thres", SCORE_THRES)), float(getattr(self, "nms_thres", NMS_THRES)),
           int(getattr(self, "detections_per_img", MAX_DETECTIONS_PER_IMAGE)), torch.cuda.current_stream(x.device).cuda_stream)
    g = cache.get(key)
    if g is None:
        from .graphs import HotPathGraph
        if len(cache) >= 4:
            cache.pop(next(iter(cache)))
        g = HotPathGraph(x.shape[2], x.detach(), b.detach(), an, im_szs, train=False, detect=True, score_thres=key[7],
                         nms_thres=key[8], detections_per_img=key[9], original_image_sizes=original_image_sizes,
                         box_format=box_format)
        g.release_inputs()
        cache[key] = g
    return _GraphPending(g.step(), x, b)


def process_detections_async(self, outputs: Dict[str, Tensor], anchors: List[Tensor],
                             im_szs: List[Tuple[int, int]],
                             original_image_sizes: Optional[Sequence[Tuple[int, int]]] = None,
                             box_format: str = "xyxy") -> PendingDetections:
    """Same arguments and side effects as :func:`process_detections`; returns a handle whose
    ``.detections()`` yields the reference's ``List[Dict]``.  ``original_image_sizes`` folds
    ``transform.postprocess``'s box resize (models.py:271) into the output write (row N2); ``box_format="xywh"``
    writes COCO boxes (row N4, see :meth:`PendingDetections.coco_results`)."""
    an, stride = _shared_anchors(anchors)
    if "cls_levels" in outputs:                       # raw per-level conv outputs
        cls_levels, box_levels = outputs.pop("cls_levels"), outputs.pop("bbox_levels")
        C = getattr(self, "num_classes", None) or cls_levels[0].shape[1] // (box_levels[0].shape[1] // 4)
        return postprocess_levels_async(cls_levels, box_levels, C, an, stride, im_szs,
                                        getattr(self, "score_thres", SCORE_THRES), getattr(self, "nms_thres", NMS_THRES),
                                        getattr(self, "detections_per_img", MAX_DETECTIONS_PER_IMAGE),
                                        getattr(self, "pre_nms_topk", None), original_image_sizes=original_image_sizes,
                                        box_format=box_format)
    class_logits = outputs.pop("cls_preds")
    bboxes = outputs.pop("bbox_preds")
    if getattr(self, "rn_graph", False) and stride == 0:
        pending = _graph_detections(self, class_logits.detach(), bboxes.detach(), an, im_szs, original_image_sizes, box_format)
        if pending is not None:
            return pending
    return postprocess_batch_async(class_logits, bboxes, an, stride, im_szs,
                                   getattr(self, "score_thres", SCORE_THRES), getattr(self, "nms_thres", NMS_THRES),
                                   getattr(self, "detections_per_img", MAX_DETECTIONS_PER_IMAGE),
                                   getattr(self, "pre_nms_topk", None), _level_offsets(self),
                                   original_image_sizes=original_image_sizes, box_format=box_format)


def process_detections(self, outputs: Dict[str, Tensor], anchors: List[Tensor],
                       im_szs: List[Tuple[int, int]],
                       original_image_sizes: Optional[Sequence[Tuple[int, int]]] = None) -> List[Dict[str, Tensor]]:
    return process_detections_async(self, outputs, anchors, im_szs, original_image_sizes).detections()
