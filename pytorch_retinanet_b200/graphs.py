"""The whole dense per-anchor path of one batch as ONE CUDA graph with two concurrent branches.

The drop-in entry points (``RetinaNetLosses.forward``, ``process_detections``) cost ~0.9 ms of host
time per batch of 16 images (torch allocations, autograd, ctypes marshalling) for ~0.6 ms of GPU
work, and they run the training half and the inference half back to back.  ``HotPathGraph`` captures
the SAME C-ABI calls (``rn_train_loss`` = matcher + loss with gradients; ``rn_postprocess``) once, for fixed
shapes and static buffers, into a CUDA graph:

    capture stream:   rn_train_loss: match -> loss (fwd + gradients) -> finalize ----.
    side stream (hi): score filter -> lazy NMS -> status --------------------------+--> join

so that a step is one ``cudaGraphLaunch`` (+ one ``rn_pack_targets`` launch for the ragged ground
truth and one 4*(N+4)-byte D2H copy of the detection counts), and the
latency-bound NMS (one CTA per image) runs in the shadow of the HBM-bound loss kernel.
Results are bit-identical to the drop-in calls (tests/test_gpu_graph.py) — the kernels and their
arguments are the same; only the launch mechanism differs.

Reference flow covered: retinanet/losses.py:113-145 (+ box_utils.py:51-80) and
retinanet/models.py:160-243, on the head outputs of retinanet/layers.py:110-115.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _native
from .box_utils import _REG_WEIGHTS_C
from .config import (FOCAL_LOSS_ALPHA, FOCAL_LOSS_GAMMA, IOU_THRESHOLDS_BACKGROUND, IOU_THRESHOLDS_FOREGROUND,
                     MAX_DETECTIONS_PER_IMAGE, NMS_THRES, SCORE_THRES, SMOOTH_L1_LOSS_BETA)
from .detections import (_FORMATS, _image_sizes_tensor, _resize_ratio_tensor, default_candidate_capacity,
                         postprocess_batch, postprocess_levels_async, slice_detections)
from .losses import _level_desc, _ptr_array

_vp = ctypes.c_void_p


class GraphStepResult:
    """Outcome of one :meth:`HotPathGraph.step`.  Tensors are views of the graph's static buffers: they stay
    valid until the next ``step`` of the same graph."""

    def __init__(self, owner: "HotPathGraph", event: Optional[torch.cuda.Event]):
        self._o = owner
        self._event = event
        self._dets = None
        self._done = None            # set by HotPathPipeline: the step ran on another stream than the consumer's

    def _consumer_waits(self) -> None:
        if self._done is not None:
            torch.cuda.current_stream(self._o.dev).wait_event(self._done)
            self._done = None

    @property
    def losses(self) -> Dict[str, Tensor]:
        """{"classification_loss", "regression_loss"} of the (global) batch — device scalars, no sync."""
        self._consumer_waits()
        t = self._o.total
        return {"classification_loss": t[0], "regression_loss": t[1]}

    @property
    def per_image(self) -> Tensor:
        """[N,3] = cls_i / max(1,F_i), reg_i / max(1,F_i), F_i."""
        self._consumer_waits()
        return self._o.per_image

    @property
    def grads(self) -> Tuple[Tensor, Tensor]:
        """d(classification_loss)/d cls_preds and d(regression_loss)/d bbox_preds, already divided by the batch size."""
        self._consumer_waits()
        return self._o.grad_cls_preds, self._o.grad_bbox_preds

    def result(self, cls_preds=None, bbox_preds=None):
        """(boxes [N,max_det,4], scores [N,max_det], labels [N,max_det] int64, counts list[int]) — waits for the
        count copy only.  ``cls_preds`` / ``bbox_preds``: the live input tensors, needed only after
        :meth:`HotPathGraph.release_inputs` for the (rare) eager re-run."""
        o = self._o
        cls_in = o.cls_preds if cls_preds is None else cls_preds
        box_in = o.bbox_preds if bbox_preds is None else bbox_preds
        if not o.detect:
            raise RuntimeError("HotPathGraph was built with detect=False")
        if self._dets is None:
            self._consumer_waits()
            self._event.synchronize()
            host = o._host.tolist()
            N = o.N
            found, capacity, fallback = host[N], host[N + 1], host[N + 2]
            if found > capacity or fallback:
                # rare: candidate pool overflow / an image needs more rounds than the lazy budget -> eager re-run of the
                # same inputs through the drop-in call, which knows how to grow the pool and switch algorithm
                kw = dict(cand_capacity=max(found, o.cap), original_image_sizes=o.original_image_sizes, box_format=o.box_format,
                          pre_nms_topk=o.topk or None)
                if cls_in is None:
                    raise RuntimeError("HotPathGraph: the eager re-run needs the input tensors (release_inputs() was called)")
                if o.levels:
                    self._dets = postprocess_levels_async(cls_in, box_in, o.C, o.anchors, o.anchor_stride, o.im_szs,
                                                          o.score_thres, o.nms_thres, o.max_det, **kw).result()
                else:
                    self._dets = postprocess_batch(cls_in, box_in, o.anchors, o.anchor_stride, o.im_szs,
                                                   o.score_thres, o.nms_thres, o.max_det,
                                                   level_offsets=o.level_offsets if o.topk else None, **kw)
            else:
                self._dets = (o.out_boxes, o.out_scores, o.out_labels, host[:N])
        return self._dets

    def detections(self) -> List[Dict[str, Tensor]]:
        """The reference's ``List[Dict]`` (models.py:236-242): three ``split_with_sizes`` calls for the batch."""
        return slice_detections(*self.result())


class HotPathGraph:
    """CUDA graph of loss (forward + gradients) and/or post-processing for fixed shapes.

    ``cls_preds [N,A,C]`` / ``bbox_preds [N,A,4]`` are the STATIC inputs: pass the tensors the head writes into
    (or copy into ``graph.cls_preds`` / ``graph.bbox_preds``); ``anchors`` is the shared ``[A,4]`` tensor of
    :class:`AnchorGenerator`.  Both may instead be LISTS of the head's raw per-level conv outputs
    ``[N, na*C, H_l, W_l]`` / ``[N, na*4, H_l, W_l]`` (row N1: no permute/cat pass); gradients then come back as lists
    in the same layout.  ``max_targets`` bounds the total number of GT boxes of a batch (static packed
    buffers).  ``global_batch`` / ``group`` / ``exchange``: image-sharded multi-GPU use — the loss is divided by the
    global batch and the 16-byte loss vector is summed over the ranks exactly as in ``ShardedRetinaNetLosses``:
    inside the captured final reduction kernel over peer-mapped memory (``exchange="peer"``, default — the graph
    launch is then the whole step), or by one NCCL ``all_reduce`` issued after the graph (``"nccl"``).
    ``fused`` (default: on whenever it applies — train and detect on ``[N,A,C]`` inputs with ``C % 4 == 0``): ONE pass
    over the logits serves both halves (``rn_train_detect``: the loss kernel is also the score filter; the NMS then runs
    behind it) instead of two concurrent branches that each stream the logits.  Same results.
    ``split`` (fused graphs): capture the front (zeroing, matcher, loss + score filter + final reduction) and the tail
    (the NMS) as TWO graphs, ``graph_front`` / ``graph_tail`` — what :class:`HotPathPipeline` launches on two streams so
    that the NMS of one batch runs under the front of the next.
    """

    def __init__(self, num_classes: int, cls_preds: Tensor, bbox_preds: Tensor, anchors: Tensor,
                 im_szs: Optional[Sequence[Tuple[int, int]]] = None, *, train: bool = True, detect: bool = True,
                 max_targets: int = 4096, global_batch: Optional[int] = None, group=None,
                 score_thres: float = SCORE_THRES, nms_thres: float = NMS_THRES,
                 detections_per_img: int = MAX_DETECTIONS_PER_IMAGE,
                 original_image_sizes: Optional[Sequence[Tuple[int, int]]] = None, box_format: str = "xyxy",
                 alpha: float = FOCAL_LOSS_ALPHA, gamma: float = FOCAL_LOSS_GAMMA, beta: float = SMOOTH_L1_LOSS_BETA,
                 match_thr: float = IOU_THRESHOLDS_FOREGROUND, back_thr: float = IOU_THRESHOLDS_BACKGROUND,
                 cand_capacity: Optional[int] = None, concurrent: bool = True, pre_nms_topk: Optional[int] = None,
                 level_offsets: Optional[Sequence[int]] = None, exchange="peer", fused: Optional[bool] = None,
                 split: bool = False, exchange_on_tail: bool = False):
        if not (train or detect):
            raise ValueError("HotPathGraph: nothing to do (train=False, detect=False)")
        lib = _native.load()
        self.levels = isinstance(cls_preds, (list, tuple))
        _native.ptr(anchors, torch.float32, "anchors")    # CUDA + dtype + contiguity, or NativeError (no CPU path)
        if self.levels:
            cls_preds, bbox_preds = list(cls_preds), list(bbox_preds)
            for t in cls_preds + bbox_preds:
                _native.ptr(t, torch.float32, "level tensor")
            self._desc, A, N = _level_desc(cls_preds, bbox_preds, num_classes)
            self._cls_ptrs, self._box_ptrs = _ptr_array(cls_preds), _ptr_array(bbox_preds)
            dev, C = cls_preds[0].device, num_classes
        else:
            _native.ptr(cls_preds, torch.float32, "cls_preds")
            _native.ptr(bbox_preds, torch.float32, "bbox_preds")
            dev = cls_preds.device
            N, A, C = cls_preds.shape
            if C != num_classes:
                raise ValueError(f"cls_preds has {C} classes, expected {num_classes}")
            if bbox_preds.shape != (N, A, 4):
                raise ValueError(f"bbox_preds must be [{N},{A},4], got {tuple(bbox_preds.shape)}")
        if anchors.dim() == 2:
            stride = 0
        elif anchors.dim() == 3 and anchors.shape[0] == N:
            stride = anchors.shape[1]
        else:
            raise ValueError("anchors must be [A,4] (shared) or [N,A,4]")
        if anchors.shape[-2] != A:
            raise ValueError(f"anchors hold {anchors.shape[-2]} rows, cls_preds {A}")
        if detect and (im_szs is None or len(im_szs) != N):
            raise ValueError("detect=True needs one (h, w) per image in im_szs")
        assert match_thr > back_thr                       # box_utils.py:66
        self.lib, self.dev, self.N, self.A, self.C = lib, dev, N, A, C
        self.train, self.detect = train, detect
        self.cls_preds, self.bbox_preds, self.anchors, self.anchor_stride = cls_preds, bbox_preds, anchors, stride
        self.im_szs = list(im_szs) if im_szs is not None else None
        self.original_image_sizes, self.box_format = original_image_sizes, box_format
        self.score_thres, self.nms_thres, self.max_det = float(score_thres), float(nms_thres), int(detections_per_img)
        self.topk = int(pre_nms_topk) if pre_nms_topk else 0        # extension: top-k per (image, pyramid level) before NMS
        self._lvl = None
        if self.topk and not self.levels:
            if level_offsets is None:
                raise ValueError("pre_nms_topk requires level_offsets (anchor offsets of the pyramid levels)")
            self.level_offsets = [int(v) for v in level_offsets]
            self._lvl = (ctypes.c_int64 * len(self.level_offsets))(*self.level_offsets)
        if self.topk and A * C >= (1 << 32):
            raise ValueError("pre_nms_topk needs A*C < 2^32")
        self.group = group
        self.world = 1
        self._xch = None
        if global_batch is not None:
            import torch.distributed as dist
            self.world = dist.get_world_size(group) if dist.is_initialized() else 1
            if self.world > 1 and train:
                from .distributed import PeerExchange, get_exchange
                self._xch = exchange if isinstance(exchange, PeerExchange) else get_exchange(group, exchange)
        self._xref = None if self._xch is None else self._xch.ref
        self.batch_div = float(global_batch if global_batch is not None else N)
        self.hp = (float(alpha), float(gamma), float(beta), float(match_thr), float(back_thr))
        self.max_targets = int(max_targets)
        if train and self.max_targets >= (1 << 20):
            raise ValueError("max_targets must stay below 2^20 (GT index field of the packed match codes)")

        f32, i32, i64 = torch.float32, torch.int32, torch.int64
        if train:
            # static packed targets: ONE byte buffer [offsets | boxes | labels] so that host-resident targets can be
            # uploaded with a single copy; the three tensors the kernels see are views of it
            cap = max(self.max_targets, 1)
            self._off_bytes = ((N + 1) * 4 + 15) // 16 * 16
            self._tgt_buf = torch.zeros((self._off_bytes + cap * 24,), dtype=torch.uint8, device=dev)
            self.gt_off = self._tgt_buf[:(N + 1) * 4].view(i32)
            self.gt_boxes = self._tgt_buf[self._off_bytes:self._off_bytes + cap * 16].view(f32).view(cap, 4)
            self.gt_labels = self._tgt_buf[self._off_bytes + cap * 16:].view(i64)
            self.gt_labels.fill_(1)
            self.codes = torch.empty((N, A), dtype=i32, device=dev)
            self.fg = torch.empty((N,), dtype=i32, device=dev)
            self.total = torch.zeros((4,), dtype=f32, device=dev)
            self.per_image = torch.zeros((N, 3), dtype=f32, device=dev)
            if self.levels:
                self.grad_cls_preds = [torch.empty_like(t) for t in cls_preds]
                self.grad_bbox_preds = [torch.empty_like(t) for t in bbox_preds]
                self._gcls_ptrs, self._gbox_ptrs = _ptr_array(self.grad_cls_preds), _ptr_array(self.grad_bbox_preds)
                self._loss_ws_bytes = lib.rn_loss_levels_workspace_bytes(N, self._desc, len(cls_preds))
            else:
                self.grad_cls_preds = torch.empty_like(cls_preds)
                self.grad_bbox_preds = torch.empty_like(bbox_preds)
                self._loss_ws_bytes = lib.rn_train_loss_workspace_bytes(N, A, C)
            self._loss_ws = torch.empty((self._loss_ws_bytes,), dtype=torch.uint8, device=dev)
        else:
            self.total = self.per_image = self.grad_cls_preds = self.grad_bbox_preds = None
        if detect:
            M = self.max_det
            self.cap = int(cand_capacity) if cand_capacity else default_candidate_capacity(N, A, C)
            self.out_boxes = torch.empty((N, M, 4), dtype=f32, device=dev)
            self.out_scores = torch.empty((N, M), dtype=f32, device=dev)
            self.out_labels = torch.empty((N, M), dtype=i64, device=dev)
            self.meta = torch.zeros((N + 4,), dtype=i32, device=dev)          # counts [N] + status [4]
            self._host = torch.empty((N + 4,), dtype=i32, pin_memory=True)
            self._hw = _image_sizes_tensor(self.im_szs, dev)
            self._ratio = _resize_ratio_tensor(self.im_szs, original_image_sizes, dev)
            self._pp_ws_bytes = (lib.rn_postprocess_levels_workspace_bytes if self.levels else
                                 lib.rn_postprocess_workspace_bytes)(N, A, C, self.cap, M)
            self._pp_ws = torch.empty((self._pp_ws_bytes,), dtype=torch.uint8, device=dev)
        self._stage = None
        self.steps_done, self.grads_taken = 0, False
        old_mode = lib.rn_loss_set_math_mode(0)               # (test hook) the fused kernel exists for the default math only
        lib.rn_loss_set_math_mode(old_mode)
        can_fuse = (not self.levels) and train and detect and old_mode == 0 and C % 4 == 0 and A * C < (1 << 32) \
            and cls_preds.data_ptr() % 16 == 0
        if fused and not can_fuse:
            raise ValueError("HotPathGraph(fused=True) needs train and detect on [N,A,C] inputs with C % 4 == 0 (default math mode)")
        self.fused = can_fuse if fused is None else bool(fused)
        self.split = bool(split)
        if self.split and not self.fused:
            raise ValueError("split=True needs a fused graph")
        # EXPERIMENTAL (not yet validated on hardware, off by default): in a split graph, leave the loss kernel's total
        # local and sum it over the ranks with the stand-alone exchange kernel at the head of the TAIL graph — the
        # lock-step wait for the slowest rank then sits on the tail stream, off the critical path of the fronts.
        self.exchange_on_tail = bool(exchange_on_tail) and self.split and self._xch is not None
        if self.fused:
            self._td_ws_bytes = lib.rn_train_detect_workspace_bytes(N, A, C, self.cap, self.max_det)
            self._td_ws = torch.empty((self._td_ws_bytes,), dtype=torch.uint8, device=dev)
        self._side = torch.cuda.Stream(device=dev, priority=-1) if (train and detect and concurrent and not self.fused) else None
        self.graph = torch.cuda.CUDAGraph()
        self._capture()

    def release_inputs(self) -> None:
        """Drops the references to the input tensors (the captured kernels keep their ADDRESSES).  For callers that
        replay the graph only while the same addresses hold live tensors of the same shape — ``RetinaNetLosses(graph=
        True)`` / ``process_detections`` in graph mode key their graph cache on exactly that — so that the caching
        allocator can recycle the blocks between steps.  The eager re-run of a candidate-pool overflow then needs the
        tensors passed to :meth:`GraphStepResult.result` again."""
        self.cls_preds = self.bbox_preds = None

    # ---- raw C-ABI launches on the CURRENT stream (the same calls the drop-in path makes) ----
    def _enqueue_train(self):
        lib, N, A, C = self.lib, self.N, self.A, self.C
        alpha, gamma, beta, match_thr, back_thr = self.hp
        if self.levels:
            self._enqueue_match()
            rc = lib.rn_loss_levels(self._cls_ptrs, self._box_ptrs, self._desc, len(self.cls_preds), self.anchors.data_ptr(),
                                    self.anchor_stride, self.gt_boxes.data_ptr(), self.gt_off.data_ptr(), self.codes.data_ptr(),
                                    self.fg.data_ptr(), N, A, C, alpha, gamma, beta, _REG_WEIGHTS_C, self.batch_div,
                                    self.per_image.data_ptr(), self.total.data_ptr(), self._gcls_ptrs, self._gbox_ptrs,
                                    self._loss_ws.data_ptr(), self._loss_ws_bytes, _native.stream_ptr(self.dev), self._xref)
            _native.check(rc, "rn_loss_levels")
            return
        rc = lib.rn_train_loss(self.cls_preds.data_ptr(), self.bbox_preds.data_ptr(), self.anchors.data_ptr(), self.anchor_stride,
                               self.gt_boxes.data_ptr(), self.gt_labels.data_ptr(), self.gt_off.data_ptr(), N,
                               self.max_targets, A, C, match_thr,
                               back_thr, alpha, gamma, beta, _REG_WEIGHTS_C, self.batch_div, self.codes.data_ptr(),
                               self.fg.data_ptr(), self.per_image.data_ptr(), self.total.data_ptr(),
                               self.grad_cls_preds.data_ptr(), self.grad_bbox_preds.data_ptr(), self._loss_ws.data_ptr(),
                               self._loss_ws_bytes, _native.stream_ptr(self.dev), self._xref)
        _native.check(rc, "rn_train_loss")

    def _enqueue_match(self):
        lib, N, A = self.lib, self.N, self.A
        match_thr, back_thr = self.hp[3], self.hp[4]
        s = _native.stream_ptr(self.dev)
        rc = lib.rn_match(self.anchors.data_ptr(), A, self.anchor_stride, self.gt_boxes.data_ptr(), self.gt_labels.data_ptr(),
                          self.gt_off.data_ptr(), N, self.max_targets, match_thr, back_thr, None, self.codes.data_ptr(),
                          self.fg.data_ptr(), s)
        _native.check(rc, "rn_match")

    def _enqueue_loss(self, want_grad: bool = True):
        """rn_loss alone on the codes of the last step (bench.py times the streaming kernel by itself with it; no
        exchange — a timing aid must not desynchronise the ranks' step counters)."""
        lib, N, A, C = self.lib, self.N, self.A, self.C
        alpha, gamma, beta = self.hp[:3]
        s = _native.stream_ptr(self.dev)
        rc = lib.rn_loss(self.cls_preds.data_ptr(), self.bbox_preds.data_ptr(), self.anchors.data_ptr(), self.anchor_stride,
                         self.gt_boxes.data_ptr(), self.gt_off.data_ptr(), self.codes.data_ptr(), self.fg.data_ptr(), N, A, C,
                         alpha, gamma, beta, _REG_WEIGHTS_C, self.batch_div, self.per_image.data_ptr(), self.total.data_ptr(),
                         self.grad_cls_preds.data_ptr() if want_grad else None,
                         self.grad_bbox_preds.data_ptr() if want_grad else None, self._loss_ws.data_ptr(),
                         self._loss_ws_bytes, s, None)
        _native.check(rc, "rn_loss")

    def _enqueue_detect(self):
        lib, N, A, C = self.lib, self.N, self.A, self.C
        meta = self.meta.data_ptr()
        if self.levels:
            rc = lib.rn_postprocess_levels(self._cls_ptrs, self._box_ptrs, self._desc, len(self.cls_preds), self.anchors.data_ptr(),
                                           self.anchor_stride, self._hw.data_ptr(), N, A, C, self.score_thres, self.nms_thres,
                                           self.max_det, _REG_WEIGHTS_C, self.topk, 0, self.cap, self.out_boxes.data_ptr(),
                                           self.out_scores.data_ptr(), self.out_labels.data_ptr(), meta, meta + 4 * N,
                                           self._pp_ws.data_ptr(), self._pp_ws_bytes, _native.stream_ptr(self.dev),
                                           None if self._ratio is None else self._ratio.data_ptr(), _FORMATS[self.box_format])
            _native.check(rc, "rn_postprocess_levels")
            return
        rc = lib.rn_postprocess(self.cls_preds.data_ptr(), self.bbox_preds.data_ptr(), self.anchors.data_ptr(),
                                self.anchor_stride, self._hw.data_ptr(), N, A, C, self.score_thres, self.nms_thres, self.max_det,
                                _REG_WEIGHTS_C, self.topk, self._lvl, (len(self.level_offsets) - 1) if self._lvl is not None else 0, 0,
                                self.cap, self.out_boxes.data_ptr(), self.out_scores.data_ptr(),
                                self.out_labels.data_ptr(), meta, meta + 4 * N, self._pp_ws.data_ptr(), self._pp_ws_bytes,
                                _native.stream_ptr(self.dev), None if self._ratio is None else self._ratio.data_ptr(),
                                _FORMATS[self.box_format])
        _native.check(rc, "rn_postprocess")

    def _enqueue_fused(self, phases: int = 7):
        """rn_train_detect: zeroing + matcher (phase 1), loss (+ gradients) that also filters the scores and finishes the
        reduction (phase 4), lazy NMS (phase 2)."""
        lib, N, A, C = self.lib, self.N, self.A, self.C
        alpha, gamma, beta, match_thr, back_thr = self.hp
        meta = self.meta.data_ptr()

        def call(phases):
            rc = lib.rn_train_detect(self.cls_preds.data_ptr(), self.bbox_preds.data_ptr(), self.anchors.data_ptr(),
                                     self.anchor_stride, self.gt_boxes.data_ptr(), self.gt_labels.data_ptr(),
                                     self.gt_off.data_ptr(), N, self.max_targets, A, C, match_thr, back_thr, alpha, gamma, beta,
                                     _REG_WEIGHTS_C, self.batch_div, self.codes.data_ptr(), self.fg.data_ptr(),
                                     self.per_image.data_ptr(), self.total.data_ptr(), self.grad_cls_preds.data_ptr(),
                                     self.grad_bbox_preds.data_ptr(), self._hw.data_ptr(), self.score_thres, self.nms_thres,
                                     self.max_det, self.topk, self._lvl,
                                     (len(self.level_offsets) - 1) if self._lvl is not None else 0, self.cap,
                                     self.out_boxes.data_ptr(), self.out_scores.data_ptr(), self.out_labels.data_ptr(), meta,
                                     meta + 4 * N, None if self._ratio is None else self._ratio.data_ptr(),
                                     _FORMATS[self.box_format], self._td_ws.data_ptr(), self._td_ws_bytes,
                                     _native.stream_ptr(self.dev), xref, phases)
            _native.check(rc, "rn_train_detect")

        xref = self._xref
        if self.exchange_on_tail and phases != 7:            # (the eager warm-up, phases 7, keeps the in-kernel exchange)
            xref = None
            if phases & 2:
                rc = lib.rn_exchange_total(self.total.data_ptr(), self._xref, _native.stream_ptr(self.dev))
                _native.check(rc, "rn_exchange_total")
        call(phases)

    def _enqueue_all(self):
        cur = torch.cuda.current_stream(self.dev)
        if self.fused:
            self._enqueue_fused()
        elif self._side is not None:
            self._side.wait_stream(cur)                   # fork
            with torch.cuda.stream(self._side):
                self._enqueue_detect()
            self._enqueue_train()
            cur.wait_stream(self._side)                   # join
        else:
            if self.detect:
                self._enqueue_detect()
            if self.train:
                self._enqueue_train()

    def _capture(self):
        with _native.on_device(self.dev):
            warm = torch.cuda.Stream(device=self.dev)
            warm.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(warm):                 # eager warm-up: module loading / func attributes before capture
                self._enqueue_all()
            torch.cuda.current_stream(self.dev).wait_stream(warm)
            torch.cuda.synchronize(self.dev)
            recorded = self.lib.rn_launch_count()
            if self.split:
                self.graph_front, self.graph_tail = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph_front):
                    self._enqueue_fused(1 | 4)
                with torch.cuda.graph(self.graph_tail):
                    self._enqueue_fused(2)
            else:
                with torch.cuda.graph(self.graph):
                    self._enqueue_all()
            self.kernel_nodes = int(self.lib.rn_launch_count() - recorded)    # kernel launches recorded into the graph(s)

    # ---- per step ----
    def load_targets(self, targets: Sequence[Dict[str, Tensor]]) -> None:
        """Packs ``targets[i]["boxes"] [G_i,4] fp32 / ["labels"] [G_i] int64`` (CUDA tensors, the reference's
        format, losses.py:126-128) into the graph's static buffers with one ``rn_pack_targets`` launch."""
        N = self.N
        if len(targets) != N:
            raise ValueError(f"{len(targets)} targets for {N} images")
        boxes = [t["boxes"] for t in targets]
        labels = [t["labels"] for t in targets]
        counts = [b.shape[0] if b.numel() else 0 for b in boxes]
        if sum(counts) > self.max_targets:
            raise ValueError(f"{sum(counts)} GT boxes exceed max_targets={self.max_targets} of this graph")
        dev = self.dev
        if sum(counts) and all(not b.is_cuda for b, c in zip(boxes, counts) if c):
            return self._load_targets_from_host(boxes, labels, counts)
        for b, l, c in zip(boxes, labels, counts):
            if c and not (b.dtype == torch.float32 and b.device == dev and b.is_contiguous() and l.dtype == torch.int64
                          and l.device == dev and l.is_contiguous() and l.numel() == c and b.dim() == 2 and b.shape[1] == 4):
                raise _native.NativeError("HotPathGraph.load_targets: boxes must be contiguous CUDA fp32 [G,4] and labels "
                                          "contiguous CUDA int64 [G] on the graph's device")
        bp = (_vp * N)(*[b.data_ptr() if c else None for b, c in zip(boxes, counts)])
        lp = (_vp * N)(*[l.data_ptr() if c else None for l, c in zip(labels, counts)])
        cnt = (ctypes.c_int32 * N)(*counts)
        with _native.on_device(dev):
            rc = self.lib.rn_pack_targets(bp, lp, cnt, N, None, self.gt_boxes.data_ptr(), self.gt_labels.data_ptr(),
                                          self.gt_off.data_ptr(), _native.stream_ptr(dev))
        _native.check(rc, "rn_pack_targets")

    def _load_targets_from_host(self, boxes, labels, counts) -> None:
        """Host-resident targets (the reference's ``collate_fn`` output): one pinned image of the static target buffer,
        one asynchronous H2D copy (no per-tensor copies, no packing launch)."""
        N, total, cap = self.N, sum(counts), max(self.max_targets, 1)
        # two pinned staging blocks alternate (allocated once): the block written now is not the one whose copy of the
        # previous step may still be in flight; its own last copy (two steps ago) is waited for, which never blocks
        # in steady state
        if self._stage is None:
            self._stage = [torch.empty((self._off_bytes + cap * 24,), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
            self._stage_ev = [None, None]
            self._stage_k = 0
        k = self._stage_k
        self._stage_k = 1 - k
        if self._stage_ev[k] is not None:
            self._stage_ev[k].synchronize()
        stage = self._stage[k]
        offs = [0]
        for c in counts:
            offs.append(offs[-1] + c)
        stage[:(N + 1) * 4].view(torch.int32).copy_(torch.tensor(offs, dtype=torch.int32))
        hb = stage[self._off_bytes:self._off_bytes + total * 16].view(torch.float32).view(total, 4)
        torch.cat([b.reshape(-1, 4).to(torch.float32) for b, c in zip(boxes, counts) if c], out=hb)
        lab0 = self._off_bytes + cap * 16
        hl = stage[lab0:lab0 + total * 8].view(torch.int64)
        live = [l.reshape(-1).to(torch.int64) for l, c in zip(labels, counts) if c]
        if sum(l.numel() for l in live) != total:
            raise ValueError("targets: number of labels does not match number of boxes")
        torch.cat(live, out=hl)
        # only the used prefix of each section matters; three slices of one pinned block, stream-ordered before the graph
        self._tgt_buf[:self._off_bytes + total * 16].copy_(stage[:self._off_bytes + total * 16], non_blocking=True)
        self._tgt_buf[lab0:lab0 + total * 8].copy_(stage[lab0:lab0 + total * 8], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        self._stage_ev[k] = ev

    def step(self, targets: Optional[Sequence[Dict[str, Tensor]]] = None) -> GraphStepResult:
        """One pass of the path over the current contents of ``cls_preds`` / ``bbox_preds``.  ``targets=None`` keeps the
        ground truth loaded by the last :meth:`load_targets`."""
        if self.train and targets is not None:
            self.load_targets(targets)
        self.steps_done += 1
        self.grads_taken = False
        if self.split:
            self.graph_front.replay()
            self.graph_tail.replay()
        else:
            self.graph.replay()
        if self.train and self.world > 1 and self._xch is None:
            import torch.distributed as dist
            dist.all_reduce(self.total, group=self.group)       # exchange="nccl": one 16-byte all-reduce after the graph
        ev = None
        if self.detect:
            self._host.copy_(self.meta, non_blocking=True)      # the single D2H copy of the path
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.dev))
        return GraphStepResult(self, ev)


class HotPathPipeline:
    """Two fused, split :class:`HotPathGraph` s on two input buffers: the fronts (zeroing, matcher, loss + score filter
    — the HBM-bound part) run back to back on one stream, the tails (the latency-bound NMS: one SM per image) on a second
    one, so the NMS of step i runs UNDER the front of step i+1.  The loss kernels themselves never overlap (two HBM-bound
    kernels only slow each other down), and the order of the steps' exchanges is the order of the fronts on every rank.
    Ordinary stream events order everything: tail(i) after front(i); front(i+2) after tail(i) (same buffer set).
    (Measured on B200, config 2: 0.432 ms per step against 0.468 ms for the single graph.  Putting the ALU-bound matcher
    of step i+1 on a third stream under the loss kernel of step i as well was measured too: 0.503 ms — its CTAs displace
    the streaming kernel's, as the single-launch fusions of round 1 already showed; not shipped.)

    ``inputs`` = two ``(cls_preds, bbox_preds)`` pairs (the static input buffers, written by the producer — e.g. the
    head — in alternation); every other argument as for :class:`HotPathGraph`.  ``step(targets)`` runs the next buffer's
    graphs and returns its :class:`GraphStepResult`; consuming a result one step late keeps the host off the critical
    path.  Results are those of the single graph (the kernels and their arguments are the same)."""

    def __init__(self, num_classes: int, inputs, anchors: Tensor, im_szs, **kw):
        if len(inputs) != 2:
            raise ValueError("HotPathPipeline takes exactly two (cls_preds, bbox_preds) input pairs")
        dev = inputs[0][0].device
        self.dev = dev
        self.front_stream = torch.cuda.Stream(device=dev)
        self.tail_stream = torch.cuda.Stream(device=dev, priority=-1)
        self.streams = (self.front_stream, self.tail_stream)
        self.graphs = [HotPathGraph(num_classes, x, b, anchors, im_szs, fused=True, split=True, **kw) for x, b in inputs]
        self.kernel_nodes = self.graphs[0].kernel_nodes
        self._tail_done = [None, None]
        self._next = 0

    def step(self, targets=None) -> GraphStepResult:
        k = self._next
        self._next = 1 - k
        g, F, T = self.graphs[k], self.front_stream, self.tail_stream
        F.wait_stream(torch.cuda.current_stream(self.dev))          # the producer wrote this buffer on the caller's stream
        if self._tail_done[k] is not None:
            F.wait_event(self._tail_done[k])                        # the tail of step i-2 still reads this buffer set
        with torch.cuda.stream(F):
            if g.train and targets is not None:
                g.load_targets(targets)
            g.steps_done += 1
            g.grads_taken = False
            g.graph_front.replay()
            if g.train and g.world > 1 and g._xch is None:          # exchange="nccl": one 16-byte all-reduce behind the front
                import torch.distributed as dist
                dist.all_reduce(g.total, group=g.group)
            front = torch.cuda.Event()
            front.record(F)
        T.wait_event(front)
        with torch.cuda.stream(T):
            g.graph_tail.replay()
            g._host.copy_(g.meta, non_blocking=True)                # the single D2H copy of the path
            done = torch.cuda.Event()
            done.record(T)
        self._tail_done[k] = done
        res = GraphStepResult(g, done)
        res._done = done
        return res
