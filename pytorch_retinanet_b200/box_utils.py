"""Drop-in box coding + Matcher (reference: retinanet/box_utils.py), CUDA only.

Same function names, argument meaning and error behaviour (``assert match_thr > back_thr``).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _native
from .config import BBOX_REG_WEIGHTS, IOU_THRESHOLDS_BACKGROUND, IOU_THRESHOLDS_FOREGROUND
from .utilities import ifnone

_REG_WEIGHTS_C = _native.host_floats(BBOX_REG_WEIGHTS)


def convert_xywh(boxes: Tensor) -> Tensor:
    """xyxy -> (cx, cy, w, h) (reference: box_utils.py:11-15).  Public helper kept for API parity; the
    kernels fuse this conversion and never call it."""
    center = (boxes[:, :2] + boxes[:, 2:]) / 2
    sizes = boxes[:, 2:] - boxes[:, :2]
    return torch.cat([center, sizes], 1)


def convert_x1y1x2y2(boxes: Tensor) -> Tensor:
    """(cx, cy, w, h) -> xyxy (reference: box_utils.py:18-22).  API-parity helper, see convert_xywh."""
    return torch.cat([boxes[:, :2] - boxes[:, 2:] / 2, boxes[:, :2] + boxes[:, 2:] / 2], 1)


def _f32c(t: Tensor) -> Tensor:
    return t.detach().to(torch.float32).contiguous()


def _boxcode(fn_name: str, a: Tensor, anchors: Tensor) -> Tensor:
    lib = _native.load()
    a32, an32 = _f32c(a).reshape(-1, 4), _f32c(anchors).reshape(-1, 4)
    if a32.shape != an32.shape:
        raise ValueError(f"{fn_name}: shapes {tuple(a.shape)} and {tuple(anchors.shape)} differ")
    out = torch.empty_like(a32)
    with _native.on_device(a32.device):
        rc = getattr(lib, fn_name)(_native.ptr(a32, what=fn_name + " input"), _native.ptr(an32, what="anchors"),
                                   a32.shape[0], _REG_WEIGHTS_C, _native.ptr(out),
                                   _native.stream_ptr(a32.device))
    _native.check(rc, fn_name)
    return out.reshape(a.shape)


def bbox_2_activ(bboxes: Tensor, anchors: Tensor) -> Tensor:
    """Regression targets of `bboxes` w.r.t. `anchors` (reference: box_utils.py:25-34)."""
    return _boxcode("rn_encode", bboxes, anchors)


def activ_2_bbox(activations: Tensor, anchors: Tensor) -> Tensor:
    """Decode model activations to boxes (reference: box_utils.py:37-48, including its
    ``sizes = a_wh * exp(dx,dy)`` quirk).  The reference divides ``activations`` in place by
    BBOX_REG_WEIGHTS (= 1.0, a numeric no-op); this implementation leaves the input untouched."""
    return _boxcode("rn_decode", activations, anchors)


class PackedTargets:
    """Ragged ground truth of a batch packed for the C ABI: boxes [sumG,4], labels [sumG], offsets [N+1].

    Host cost matters here (the GPU part of a step is < 1 ms): one ``torch.cat`` per field, offsets
    computed from shapes (no device sync) and shipped through pinned memory."""

    def __init__(self, boxes: Sequence[Tensor], labels: Optional[Sequence[Tensor]], device: torch.device,
                 ratios_hw: Optional[Sequence[Tuple[float, float]]] = None):
        counts = [b.shape[0] if b.numel() else 0 for b in boxes]       # numel()==0 -> "no targets" (box_utils.py:70)
        self.num_images = len(counts)
        self.total = sum(counts)
        self.counts = counts
        device = torch.device(device)
        if device.type == "cuda" and self.total and all(not b.is_cuda for b, c in zip(boxes, counts) if c) and \
                self._pack_from_host(boxes, labels, counts, device, ratios_hw):
            return
        if device.type == "cuda" and self._pack_native(boxes, labels, counts, device, ratios_hw):
            return
        # generic path (dtype/device conversions needed, or host-side use in the CPU tests): torch ops
        offs = [0]
        for c in counts:
            offs.append(offs[-1] + c)
        with torch.no_grad():
            if self.total:
                live = [b.reshape(-1, 4) for b, c in zip(boxes, counts) if c]
                if ratios_hw is not None:
                    live = [b.to(torch.float32) * b.new_tensor([rw, rh, rw, rh], dtype=torch.float32)
                            for b, (rh, rw) in zip(live, [r for r, c in zip(ratios_hw, counts) if c])]
                bx = live[0] if len(live) == 1 else torch.cat(live)
                if bx.dtype != torch.float32 or bx.device != device:
                    bx = bx.to(device=device, dtype=torch.float32)
                self.boxes = bx.reshape(-1, 4).contiguous()
            else:
                self.boxes = torch.zeros((1, 4), dtype=torch.float32, device=device)
            if labels is not None:
                if self.total:
                    ll = [l for l, c in zip(labels, counts) if c]
                    lb = ll[0] if len(ll) == 1 else torch.cat(ll)
                    if lb.dtype != torch.int64 or lb.device != device:
                        lb = lb.to(device=device, dtype=torch.int64)
                    self.labels = lb.reshape(-1).contiguous()
                    if self.labels.shape[0] != self.total:
                        raise ValueError("targets: number of labels does not match number of boxes")
                else:
                    self.labels = torch.zeros((1,), dtype=torch.int64, device=device)
            else:
                self.labels = None
            # <= a few hundred bytes: the driver embeds a pageable copy of this size in the command stream and
            # returns without waiting for the GPU, which is cheaper than allocating pinned staging memory
            self.offsets = torch.tensor(offs, dtype=torch.int32, device=device)

    def _pack_from_host(self, boxes, labels, counts, device, ratios_hw) -> bool:
        """Targets still on the HOST (what ``collate_fn`` hands over, utils/detection_utils.py:7-9): instead of one
        small pageable host->device copy per tensor (2N per batch), offsets, boxes and labels are laid out in ONE
        pinned staging buffer and shipped with ONE asynchronous copy; the packed tensors are views of the device
        copy.  The optional box resize is the same fp32 multiply torchvision's ``resize_boxes`` performs."""
        N, total = len(counts), self.total
        if labels is not None:
            for l, c in zip(labels, counts):
                if l.numel() != c:
                    raise ValueError("targets: number of labels does not match number of boxes")
        off_bytes = ((N + 1) * 4 + 15) // 16 * 16
        box_bytes = total * 16
        lab_bytes = total * 8 if labels is not None else 0
        stage = torch.empty((off_bytes + box_bytes + lab_bytes,), dtype=torch.uint8, pin_memory=True)
        offs = [0]
        for c in counts:
            offs.append(offs[-1] + c)
        stage[:(N + 1) * 4].view(torch.int32).copy_(torch.tensor(offs, dtype=torch.int32))
        hb = stage[off_bytes:off_bytes + box_bytes].view(torch.float32).view(total, 4)
        live = [b.reshape(-1, 4).to(torch.float32) for b, c in zip(boxes, counts) if c]
        if ratios_hw is not None:
            live = [b * torch.tensor([rw, rh, rw, rh], dtype=torch.float32)
                    for b, (rh, rw) in zip(live, [r for r, c in zip(ratios_hw, counts) if c])]
        torch.cat(live, out=hb)
        if labels is not None:
            hl = stage[off_bytes + box_bytes:].view(torch.int64)
            torch.cat([l.reshape(-1).to(torch.int64) for l, c in zip(labels, counts) if c], out=hl)
        dev_buf = stage.to(device, non_blocking=True)               # the one H2D copy
        self._staging = stage                                       # keep the pinned block alive until the copy ran
        self.offsets = dev_buf[:(N + 1) * 4].view(torch.int32)
        self.boxes = dev_buf[off_bytes:off_bytes + box_bytes].view(torch.float32).view(total, 4)
        self.labels = dev_buf[off_bytes + box_bytes:].view(torch.int64) if labels is not None else None
        return True

    def _pack_native(self, boxes, labels, counts, device, ratios_hw) -> bool:
        """One ``rn_pack_targets`` launch (row N3) when every tensor already is a contiguous fp32 / int64 tensor
        on ``device``; returns False to let the torch path handle conversions."""
        for b, c in zip(boxes, counts):
            if c and not (b.dtype == torch.float32 and b.device == device and b.is_contiguous() and b.dim() == 2
                          and b.shape[1] == 4 and b.data_ptr() % 16 == 0):
                return False
        if labels is not None:
            for l, c in zip(labels, counts):
                if l.numel() != c:
                    raise ValueError("targets: number of labels does not match number of boxes")
                if c and not (l.dtype == torch.int64 and l.device == device and l.is_contiguous()):
                    return False
        N, total = len(counts), self.total
        self.boxes = torch.empty((max(total, 1), 4), dtype=torch.float32, device=device)
        self.labels = torch.empty((max(total, 1),), dtype=torch.int64, device=device) if labels is not None else None
        self.offsets = torch.empty((N + 1,), dtype=torch.int32, device=device)
        vp = ctypes.c_void_p
        bp = (vp * N)(*[b.data_ptr() if c else None for b, c in zip(boxes, counts)])
        lp = (vp * N)(*[l.data_ptr() if c else None for l, c in zip(labels, counts)]) if labels is not None else None
        cnt = (ctypes.c_int32 * N)(*counts)
        rat = None
        if ratios_hw is not None:
            flat = [float(v) for r in ratios_hw for v in r]
            rat = (ctypes.c_float * len(flat))(*flat)
        with _native.on_device(device):
            rc = _native.load().rn_pack_targets(bp, lp, cnt, N, rat, self.boxes.data_ptr(),
                                                None if self.labels is None else self.labels.data_ptr(),
                                                self.offsets.data_ptr(), _native.stream_ptr(device))
        _native.check(rc, "rn_pack_targets")
        return True


def match_batch(anchors: Tensor, anchor_stride: int, packed: PackedTargets, num_anchors: int,
                match_thr: float, back_thr: float, want_matches: bool, want_codes: bool
                ) -> Tuple[Optional[Tensor], Optional[Tensor], Optional[Tensor]]:
    """Runs ``rn_match`` for a batch.  Returns (matches int64 [N,A] | None, codes int32 [N,A] | None, fg_count [N] | None)."""
    lib = _native.load()
    dev = anchors.device
    N, A = packed.num_images, num_anchors
    matches = torch.empty((N, A), dtype=torch.int64, device=dev) if want_matches else None
    codes = torch.empty((N, A), dtype=torch.int32, device=dev) if want_codes else None
    fg = torch.empty((N,), dtype=torch.int32, device=dev) if want_codes else None   # zeroed by rn_match
    with _native.on_device(dev):
        rc = lib.rn_match(_native.ptr(anchors, torch.float32, "anchors"), A, anchor_stride,
                          _native.ptr(packed.boxes, torch.float32, "target boxes"),
                          _native.ptr(packed.labels, torch.int64, "target labels") if want_codes else None,
                          _native.ptr(packed.offsets), N, packed.total, float(match_thr), float(back_thr),
                          _native.ptr(matches), _native.ptr(codes), _native.ptr(fg), _native.stream_ptr(dev))
    _native.check(rc, "rn_match")
    return matches, codes, fg


def matcher(anchors: Tensor, targets: Tensor, match_thr: float = None, back_thr: float = None) -> Tensor:
    """Match `anchors` to `targets`: -1 = background, -2 = ignore, g >= 0 = GT index
    (reference: box_utils.py:51-80).  Fused IoU + argmax + thresholds; no [G,A] matrix."""
    match_thr = ifnone(match_thr, IOU_THRESHOLDS_FOREGROUND)
    back_thr = ifnone(back_thr, IOU_THRESHOLDS_BACKGROUND)
    assert match_thr > back_thr
    an = _f32c(anchors).reshape(-1, 4)
    if not an.is_cuda:
        raise _native.NativeError("retinanet_b200.matcher: anchors must be a CUDA tensor; there is no CPU path")
    packed = PackedTargets([targets.to(an.device)], None, an.device)
    m, _, _ = match_batch(an, 0, packed, an.shape[0], match_thr, back_thr, True, False)
    return m[0]
