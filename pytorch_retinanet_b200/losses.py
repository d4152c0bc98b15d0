"""Drop-in ``RetinaNetLosses`` (reference: retinanet/losses.py) — fused CUDA matcher + loss.

``forward(targets, head_outputs, anchors)`` keeps the reference signature and returns the same
dict (``classification_loss``, ``regression_loss``); both are differentiable w.r.t. ``cls_preds``
and ``bbox_preds``.  The whole batch is handled by two kernels (rn_match, rn_loss) plus a tiny
fixed-order reduction; gradients are produced in the same pass as the loss, and ``backward`` only
rescales them by autograd's grad_output (a no-op launch when it is 1).
No host synchronisation happens anywhere in forward or backward.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import Tensor, nn

from . import _native
from .box_utils import _REG_WEIGHTS_C, PackedTargets, match_batch
from .config import (BBOX_REG_WEIGHTS, FOCAL_LOSS_ALPHA, FOCAL_LOSS_GAMMA, IOU_THRESHOLDS_BACKGROUND,
                     IOU_THRESHOLDS_FOREGROUND, SMOOTH_L1_LOSS_BETA)


def _shared_anchors(anchors: Sequence[Tensor]) -> Tuple[Tensor, int]:
    """(anchor tensor, image stride).  The same tensor object for every image (what our
    AnchorGenerator returns) -> shared [A,4], stride 0; otherwise stacked [N,A,4], stride A."""
    if isinstance(anchors, Tensor):
        a = anchors.detach().to(torch.float32).contiguous()
        return (a, 0) if a.dim() == 2 else (a, a.shape[1])
    first = anchors[0]
    if all(a is first for a in anchors):
        return first.detach().to(torch.float32).contiguous(), 0
    st = torch.stack([a.detach().to(torch.float32) for a in anchors]).contiguous()
    return st, st.shape[1]


def fused_loss_forward(cls_preds: Tensor, bbox_preds: Tensor, anchors: Tensor, anchor_stride: int,
                       packed: PackedTargets, alpha: float, gamma: float, beta: float, match_thr: float,
                       back_thr: float, batch_div: float, want_grad: bool):
    """Launches rn_match + rn_loss.  Returns (out_total [4], out_image [N,3], grad_logits|None, grad_bbox|None, codes)."""
    lib = _native.load()
    dev = cls_preds.device
    N, A, C = cls_preds.shape
    if bbox_preds.shape != (N, A, 4):
        raise ValueError(f"bbox_preds must be [{N},{A},4], got {tuple(bbox_preds.shape)}")
    if packed.num_images != N:
        raise ValueError(f"{packed.num_images} targets for {N} images")
    x = cls_preds.detach()
    b = bbox_preds.detach()
    x = x if (x.dtype == torch.float32 and x.is_contiguous()) else x.to(torch.float32).contiguous()
    b = b if (b.dtype == torch.float32 and b.is_contiguous()) else b.to(torch.float32).contiguous()
    _, codes, fg = match_batch(anchors, anchor_stride, packed, A, match_thr, back_thr, False, True)
    out_total = torch.empty((4,), dtype=torch.float32, device=dev)
    out_image = torch.empty((N, 3), dtype=torch.float32, device=dev)
    gl = torch.empty_like(x) if want_grad else None
    gb = torch.empty_like(b) if want_grad else None
    ws_bytes = lib.rn_loss_workspace_bytes(N, A, C)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.rn_loss(_native.ptr(x, torch.float32, "cls_preds"), _native.ptr(b, torch.float32, "bbox_preds"),
                         _native.ptr(anchors, torch.float32, "anchors"), anchor_stride,
                         _native.ptr(packed.boxes), _native.ptr(packed.offsets), _native.ptr(codes), _native.ptr(fg),
                         N, A, C, float(alpha), float(gamma), float(beta), _REG_WEIGHTS_C,
                         float(batch_div), _native.ptr(out_image), _native.ptr(out_total), _native.ptr(gl),
                         _native.ptr(gb), _native.ptr(ws), ws_bytes, _native.stream_ptr(dev))
    _native.check(rc, "rn_loss")
    return out_total, out_image, gl, gb, codes


class _FusedRetinaNetLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cls_preds, bbox_preds, anchors, anchor_stride, packed, hp):
        want = cls_preds.requires_grad or bbox_preds.requires_grad
        total, image, gl, gb, _ = fused_loss_forward(cls_preds, bbox_preds, anchors, anchor_stride, packed,
                                                     hp["alpha"], hp["gamma"], hp["beta"], hp["match_thr"],
                                                     hp["back_thr"], hp["batch_div"], want)
        group = hp.get("all_reduce_group", False)
        if group is not False:
            # image-sharded batch: the ONE collective of the path — 16 bytes over NCCL/NVLink.  The local
            # gradients are already scaled by 1/N_global, so backward needs no communication.
            import torch.distributed as dist
            dist.all_reduce(total, group=group)
        ctx.grads = (gl, gb)
        ctx.in_dtypes = (cls_preds.dtype, bbox_preds.dtype)
        ctx.mark_non_differentiable(image, total)
        return total[0], total[1], image, total

    @staticmethod
    def backward(ctx, g_cls, g_reg, _g_image, _g_total):
        gl, gb = ctx.grads
        ctx.grads = (None, None)
        if gl is None:
            return None, None, None, None, None, None
        lib = _native.load()
        outs = []
        for buf, g, dt in ((gl, g_cls, ctx.in_dtypes[0]), (gb, g_reg, ctx.in_dtypes[1])):
            if g is None:
                outs.append(None)
                continue
            gs = g.detach().to(device=buf.device, dtype=torch.float32).contiguous()
            with torch.cuda.device(buf.device):
                rc = lib.rn_scale_by_device_scalar(_native.ptr(buf), buf.numel(), _native.ptr(gs),
                                                   _native.stream_ptr(buf.device))
            _native.check(rc, "rn_scale_by_device_scalar")
            outs.append(buf if dt == torch.float32 else buf.to(dt))
        return outs[0], outs[1], None, None, None, None


class _DenseLoss(torch.autograd.Function):
    """sum-reduced element-wise loss with its gradient produced in the same pass (focal or smooth-L1)."""

    @staticmethod
    def forward(ctx, kind, x, t, p0, p1):
        lib = _native.load()
        xc = x.detach().to(torch.float32).contiguous()
        tc = t.detach().to(torch.float32).contiguous()
        if xc.shape != tc.shape:
            raise ValueError(f"{kind}: input {tuple(x.shape)} and target {tuple(t.shape)} differ in shape")
        dev = xc.device
        out = torch.empty((1,), dtype=torch.float32, device=dev)
        grad = torch.empty_like(xc) if x.requires_grad else None
        nb = lib.rn_dense_loss_workspace_bytes()
        ws = torch.empty((nb,), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            if kind == "focal":
                rc = lib.rn_focal_loss_dense(_native.ptr(xc, what="clas_pred"), _native.ptr(tc, what="clas_tgt"), xc.numel(),
                                             float(p0), float(p1), _native.ptr(out), _native.ptr(grad), _native.ptr(ws), nb,
                                             _native.stream_ptr(dev))
            else:
                rc = lib.rn_smooth_l1_dense(_native.ptr(xc, what="input"), _native.ptr(tc, what="target"), xc.numel(),
                                            float(p0), _native.ptr(out), _native.ptr(grad), _native.ptr(ws), nb,
                                            _native.stream_ptr(dev))
        _native.check(rc, kind)
        ctx.grad = grad
        ctx.in_dtype = x.dtype
        return out[0]

    @staticmethod
    def backward(ctx, g):
        if ctx.grad is None:
            return None, None, None, None, None
        gx = ctx.grad * g                      # tiny API-parity path; the fused path never comes here
        return None, gx.to(ctx.in_dtype), None, None, None


class RetinaNetLosses(nn.Module):
    """Reference: retinanet/losses.py:11-145 (hyper-parameters read from config at construction)."""

    def __init__(self, num_classes: int) -> None:
        super().__init__()
        self.n_c = num_classes
        self.alpha = FOCAL_LOSS_ALPHA
        self.gamma = FOCAL_LOSS_GAMMA
        self.beta = SMOOTH_L1_LOSS_BETA

    def _hp(self, batch_div: float) -> dict:
        return {"alpha": self.alpha, "gamma": self.gamma, "beta": self.beta,
                "match_thr": IOU_THRESHOLDS_FOREGROUND, "back_thr": IOU_THRESHOLDS_BACKGROUND,
                "batch_div": float(batch_div)}

    def smooth_l1_loss(self, input: Tensor, target: Tensor) -> Tensor:
        """Sum-reduced smooth-L1 (reference: losses.py:19-27); differentiable w.r.t. ``input``."""
        return _DenseLoss.apply("smooth_l1", input, target, self.beta, 0.0)

    def focal_loss(self, clas_pred: Tensor, clas_tgt: Tensor) -> Tensor:
        """Sum-reduced sigmoid focal loss on dense float targets (reference: losses.py:29-47, weights from
        detached probabilities, alpha applied inverted, NO +1 shift — that lives in calc_loss)."""
        return _DenseLoss.apply("focal", clas_pred, clas_tgt, self.alpha, self.gamma)

    def calc_loss(self, anchors: Tensor, clas_pred: Tensor, bbox_pred: Tensor, clas_tgt: Tensor,
                  bbox_tgt: Tensor) -> Tuple[Tensor, Tensor]:
        """Loss of ONE image, returned as (bb_loss, clas_loss) like the reference (losses.py:49-111)."""
        if clas_pred.shape[-1] != self.n_c:
            raise ValueError(f"clas_pred has {clas_pred.shape[-1]} classes, expected {self.n_c}")
        an = anchors.detach().to(torch.float32).contiguous()
        packed = PackedTargets([bbox_tgt], [clas_tgt], clas_pred.device)
        c, r, _, _ = _FusedRetinaNetLoss.apply(clas_pred[None], bbox_pred[None], an, 0, packed, self._hp(1.0))
        return r, c

    def forward(self, targets: List[Dict[str, Tensor]], head_outputs: Dict[str, Tensor],
                anchors: List[Tensor]) -> Dict[str, Tensor]:
        clas_preds, bbox_preds = head_outputs["cls_preds"], head_outputs["bbox_preds"]
        if clas_preds.shape[-1] != self.n_c:
            raise ValueError(f"cls_preds has {clas_preds.shape[-1]} classes, expected {self.n_c}")
        an, stride = _shared_anchors(anchors)
        packed = PackedTargets([t["boxes"] for t in targets], [t["labels"] for t in targets], clas_preds.device)
        c, r, image, _ = _FusedRetinaNetLoss.apply(clas_preds, bbox_preds, an, stride, packed, self._hp(len(targets)))
        self.last_per_image = image   # [N,3]: cls_i, reg_i, F_i (device tensor, no sync)
        return {"classification_loss": c, "regression_loss": r}
