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
from .config import (FOCAL_LOSS_ALPHA, FOCAL_LOSS_GAMMA, IOU_THRESHOLDS_BACKGROUND,
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


_SCRATCH: Dict[tuple, Tensor] = {}


def fused_loss_forward(cls_preds: Tensor, bbox_preds: Tensor, anchors: Tensor, anchor_stride: int,
                       packed: PackedTargets, alpha: float, gamma: float, beta: float, match_thr: float,
                       back_thr: float, batch_div: float, want_grad: bool, exchange=None):
    """One C call (rn_train_loss): matcher, then loss + gradients + final reduction (+ the peer-memory exchange of the
    16-byte loss vector when ``exchange`` is a :class:`distributed.PeerExchange`).  Returns (out_total [4],
    out_image [N,3], grad_logits|None, grad_bbox|None, codes)."""
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
    assert match_thr > back_thr                    # box_utils.py:66
    # scratch of the call (match codes, foreground counts, partial sums): ONE cached block per (device, stream, shape) —
    # stream-ordered reuse, like the post-processing workspace; the returned `codes` view is valid until the next call
    ws_bytes = lib.rn_train_loss_workspace_bytes(N, A, C)
    code_bytes = (N * A * 4 + 255) // 256 * 256
    fg_bytes = (N * 4 + 255) // 256 * 256
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream, N, A, C)
    scratch = _SCRATCH.get(key)
    if scratch is None:
        if len(_SCRATCH) > 8:
            _SCRATCH.clear()
        scratch = _SCRATCH[key] = torch.empty((code_bytes + fg_bytes + ws_bytes,), dtype=torch.uint8, device=dev)
    codes = scratch[:N * A * 4].view(torch.int32).view(N, A)
    fg = scratch[code_bytes:code_bytes + N * 4].view(torch.int32)
    ws = scratch[code_bytes + fg_bytes:]
    out = torch.empty((4 + 3 * N,), dtype=torch.float32, device=dev)      # results handed to the caller: fresh per call
    out_total, out_image = out[:4], out[4:].view(N, 3)
    gl = torch.empty_like(x) if want_grad else None
    gb = torch.empty_like(b) if want_grad else None
    with _native.on_device(dev):
        rc = lib.rn_train_loss(_native.ptr(x, torch.float32, "cls_preds"), _native.ptr(b, torch.float32, "bbox_preds"),
                               _native.ptr(anchors, torch.float32, "anchors"), anchor_stride,
                               _native.ptr(packed.boxes, torch.float32, "target boxes"),
                               _native.ptr(packed.labels, torch.int64, "target labels"), _native.ptr(packed.offsets),
                               N, packed.total, A, C, float(match_thr), float(back_thr), float(alpha), float(gamma),
                               float(beta), _REG_WEIGHTS_C, float(batch_div), _native.ptr(codes), _native.ptr(fg),
                               _native.ptr(out_image), _native.ptr(out_total), _native.ptr(gl), _native.ptr(gb),
                               _native.ptr(ws), ws_bytes, _native.stream_ptr(dev),
                               None if exchange is None else exchange.ref)
    _native.check(rc, "rn_train_loss")
    return out_total, out_image, gl, gb, codes


def _scale_in_place(buf: Tensor, g: Tensor) -> None:
    """buf *= g (device scalar) with one launch whose blocks exit at once when g == 1 (the usual case)."""
    if buf.numel() == 0:
        return
    gs = g if (g.dtype == torch.float32 and g.device == buf.device and not g.requires_grad) else \
        g.detach().to(device=buf.device, dtype=torch.float32).contiguous()
    with _native.on_device(buf.device):
        rc = _native.load().rn_scale_by_device_scalar(_native.ptr(buf), buf.numel(), _native.ptr(gs),
                                                      _native.stream_ptr(buf.device))
    _native.check(rc, "rn_scale_by_device_scalar")


class _FusedRetinaNetLoss(torch.autograd.Function):
    """The gradient buffers written by the forward kernel are handed to autograd by ``backward`` (scaled in place, no
    copy of the 1 GB tensors).  A buffer that an earlier ``backward`` already gave away (``retain_graph=True`` followed
    by a second backward, or the two losses back-propagated one after the other) is recomputed from the saved inputs
    with one more pass of the forward kernel, so every autograd usage the reference supports gives the same numbers."""

    @staticmethod
    def forward(ctx, cls_preds, bbox_preds, anchors, anchor_stride, packed, hp):
        want = cls_preds.requires_grad or bbox_preds.requires_grad
        if cls_preds.shape[0] == 0:      # an empty shard of an image-sharded batch still joins the all-reduce
            total = torch.zeros((4,), dtype=torch.float32, device=cls_preds.device)
            image = torch.zeros((0, 3), dtype=torch.float32, device=cls_preds.device)
            gl = torch.zeros_like(cls_preds, dtype=torch.float32) if want else None
            gb = torch.zeros_like(bbox_preds, dtype=torch.float32) if want else None
        else:
            total, image, gl, gb, _ = fused_loss_forward(cls_preds, bbox_preds, anchors, anchor_stride, packed,
                                                         hp["alpha"], hp["gamma"], hp["beta"], hp["match_thr"],
                                                         hp["back_thr"], hp["batch_div"], want, hp.get("exchange"))
        total = _exchange_total(total, hp, in_kernel=cls_preds.shape[0] > 0)
        ctx.set_materialize_grads(False)          # a loss that is not back-propagated arrives as None, not as zeros
        ctx.grads = (gl, gb)
        ctx.in_dtypes = (cls_preds.dtype, bbox_preds.dtype)
        if want:
            ctx.save_for_backward(cls_preds, bbox_preds)
            ctx.recompute = (anchors, anchor_stride, packed, hp)
        ctx.mark_non_differentiable(image, total)
        return total[0], total[1], image, total

    @staticmethod
    def backward(ctx, g_cls, g_reg, _g_image, _g_total):
        need = (g_cls is not None and ctx.needs_input_grad[0], g_reg is not None and ctx.needs_input_grad[1])
        if not any(need):
            return None, None, None, None, None, None
        bufs = list(ctx.grads)
        if any(n and b is None for n, b in zip(need, bufs)):
            cls_preds, bbox_preds = ctx.saved_tensors
            anchors, anchor_stride, packed, hp = ctx.recompute
            if cls_preds.shape[0] == 0:
                fresh = (torch.zeros_like(cls_preds, dtype=torch.float32), torch.zeros_like(bbox_preds, dtype=torch.float32))
            else:
                fresh = fused_loss_forward(cls_preds, bbox_preds, anchors, anchor_stride, packed, hp["alpha"], hp["gamma"],
                                           hp["beta"], hp["match_thr"], hp["back_thr"], hp["batch_div"], True)[2:4]
            bufs = [b if b is not None else f for b, f in zip(bufs, fresh)]
        outs = []
        for i, g in enumerate((g_cls, g_reg)):
            if not need[i]:
                outs.append(None)
                continue
            buf, bufs[i] = bufs[i], None              # handed over: autograd may take the tensor without a copy
            _scale_in_place(buf, g)
            outs.append(buf if ctx.in_dtypes[i] == torch.float32 else buf.to(ctx.in_dtypes[i]))
        ctx.grads = tuple(bufs)
        return outs[0], outs[1], None, None, None, None


def _exchange_total(total: Tensor, hp: dict, in_kernel: bool = True) -> Tensor:
    """Image-sharded batch: the ONE exchange of the path — the sum over the ranks of the 16-byte vector
    [cls, reg, sum F, N_local].  With ``hp["exchange"]`` (a :class:`distributed.PeerExchange`) the loss's final
    reduction kernel already did it over peer memory (``in_kernel``; an empty shard, which launches no kernel, joins
    through :meth:`PeerExchange.exchange`); with ``hp["all_reduce_group"]`` it is one ``all_reduce`` (NCCL, or gloo in
    the CPU tests).  The local gradients are already scaled by 1/N_global, so backward needs no communication."""
    x = hp.get("exchange")
    if x is not None:
        if not in_kernel:
            x.exchange(total)
    else:
        group = hp.get("all_reduce_group", False)
        if group is not False:
            import torch.distributed as dist
            dist.all_reduce(total, group=group)
    scale = hp.get("total_scale", 1.0)
    if scale != 1.0:
        total[:2] *= scale
    return total


def _level_desc(cls_levels: Sequence[Tensor], box_levels: Sequence[Tensor], C: int):
    """(ctypes level_desc [L][3] = H,W,na ; total anchors A ; N) after validating the per-level NCHW tensors."""
    import ctypes
    if len(cls_levels) != len(box_levels) or not len(cls_levels):
        raise ValueError("cls_levels and bbox_levels must be non-empty lists of equal length")
    N = cls_levels[0].shape[0]
    desc, A = [], 0
    for x, b in zip(cls_levels, box_levels):
        if x.dim() != 4 or b.dim() != 4 or x.shape[0] != N or b.shape[0] != N or x.shape[-2:] != b.shape[-2:]:
            raise ValueError(f"level tensors must be [N, na*C, H, W] / [N, na*4, H, W]; got {tuple(x.shape)} / {tuple(b.shape)}")
        na = b.shape[1] // 4
        if b.shape[1] != na * 4 or x.shape[1] != na * C:
            raise ValueError(f"level channels {x.shape[1]} / {b.shape[1]} do not match na*C / na*4 with C={C}")
        H, W = int(x.shape[2]), int(x.shape[3])
        desc += [H, W, na]
        A += H * W * na
    return (ctypes.c_int32 * len(desc))(*desc), A, N


def _ptr_array(tensors: Sequence[Optional[Tensor]]):
    import ctypes
    return (ctypes.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])


def _f32_contig(t: Tensor) -> Tensor:
    t = t.detach()
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.to(torch.float32).contiguous()


def fused_loss_forward_levels(cls_levels, box_levels, anchors: Tensor, anchor_stride: int, packed: PackedTargets, C: int,
                              alpha, gamma, beta, match_thr, back_thr, batch_div, want_grad: bool, exchange=None):
    """rn_match + rn_loss_levels on the RAW per-level conv outputs: no permute/cat pass."""
    lib = _native.load()
    xs, bs = [_f32_contig(t) for t in cls_levels], [_f32_contig(t) for t in box_levels]
    desc, A, N = _level_desc(xs, bs, C)
    dev = xs[0].device
    if packed.num_images != N:
        raise ValueError(f"{packed.num_images} targets for {N} images")
    if anchors.shape[-2] != A:
        raise ValueError(f"anchors hold {anchors.shape[-2]} rows, the levels {A}")
    _, codes, fg = match_batch(anchors, anchor_stride, packed, A, match_thr, back_thr, False, True)
    out_total = torch.empty((4,), dtype=torch.float32, device=dev)
    out_image = torch.empty((N, 3), dtype=torch.float32, device=dev)
    gxs = [torch.empty_like(t) for t in xs] if want_grad else None
    gbs = [torch.empty_like(t) for t in bs] if want_grad else None
    L = len(xs)
    ws_bytes = lib.rn_loss_levels_workspace_bytes(N, desc, L)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    with _native.on_device(dev):
        rc = lib.rn_loss_levels(_ptr_array(xs), _ptr_array(bs), desc, L, _native.ptr(anchors, torch.float32, "anchors"),
                                anchor_stride, _native.ptr(packed.boxes), _native.ptr(packed.offsets), _native.ptr(codes),
                                _native.ptr(fg), N, A, C, float(alpha), float(gamma), float(beta), _REG_WEIGHTS_C,
                                float(batch_div), _native.ptr(out_image), _native.ptr(out_total),
                                _ptr_array(gxs) if want_grad else None, _ptr_array(gbs) if want_grad else None,
                                _native.ptr(ws), ws_bytes, _native.stream_ptr(dev),
                                None if exchange is None else exchange.ref)
    _native.check(rc, "rn_loss_levels")
    return out_total, out_image, gxs, gbs


class _FusedRetinaNetLossLevels(torch.autograd.Function):
    """inputs: anchors, anchor_stride, packed, hp, C, L, then L class-level tensors and L box-level tensors.
    Gradient buffers are handed over / recomputed exactly as in :class:`_FusedRetinaNetLoss`."""

    @staticmethod
    def forward(ctx, anchors, anchor_stride, packed, hp, C, L, *levels):
        cls_levels, box_levels = levels[:L], levels[L:]
        want = any(t.requires_grad for t in levels)
        total, image, gxs, gbs = fused_loss_forward_levels(cls_levels, box_levels, anchors, anchor_stride, packed, C,
                                                           hp["alpha"], hp["gamma"], hp["beta"], hp["match_thr"],
                                                           hp["back_thr"], hp["batch_div"], want, hp.get("exchange"))
        total = _exchange_total(total, hp)
        ctx.set_materialize_grads(False)
        ctx.grads = (gxs, gbs)
        ctx.L = L
        ctx.in_dtypes = [t.dtype for t in levels]
        if want:
            ctx.save_for_backward(*levels)
            ctx.recompute = (anchors, anchor_stride, packed, hp, C)
        ctx.mark_non_differentiable(image, total)
        return total[0], total[1], image, total

    @staticmethod
    def backward(ctx, g_cls, g_reg, _g_image, _g_total):
        L = ctx.L
        need = (g_cls is not None and any(ctx.needs_input_grad[6:6 + L]),
                g_reg is not None and any(ctx.needs_input_grad[6 + L:6 + 2 * L]))
        if not any(need):
            return (None,) * (6 + 2 * L)
        bufs = list(ctx.grads)
        if any(n and b is None for n, b in zip(need, bufs)):
            levels = ctx.saved_tensors
            anchors, anchor_stride, packed, hp, C = ctx.recompute
            fresh = fused_loss_forward_levels(levels[:L], levels[L:], anchors, anchor_stride, packed, C, hp["alpha"],
                                              hp["gamma"], hp["beta"], hp["match_thr"], hp["back_thr"], hp["batch_div"],
                                              True)[2:4]
            bufs = [b if b is not None else f for b, f in zip(bufs, fresh)]
        outs = []
        for i, g in enumerate((g_cls, g_reg)):
            if not need[i]:
                outs.extend([None] * L)
                continue
            mine, bufs[i] = bufs[i], None
            for buf in mine:
                _scale_in_place(buf, g)
            outs.extend(mine)
        ctx.grads = tuple(bufs)
        outs = [o if (o is None or dt == torch.float32) else o.to(dt) for o, dt in zip(outs, ctx.in_dtypes)]
        return (None,) * 6 + tuple(outs)


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
        with _native.on_device(dev):
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


class _GraphLoss(torch.autograd.Function):
    """``RetinaNetLosses(graph=True)``: the forward is one replay of a cached :class:`graphs.HotPathGraph` (train half
    only) whose kernels read the head outputs in place; backward hands out the graph's static gradient buffers."""

    @staticmethod
    def forward(ctx, cls_preds, bbox_preds, graph, targets):
        graph.step(targets)
        ctx.graph = graph
        ctx.step_id = graph.steps_done
        ctx.set_materialize_grads(False)
        ctx.in_dtypes = (cls_preds.dtype, bbox_preds.dtype)
        out = torch.cat([graph.total, graph.per_image.reshape(-1)])     # results leave the static buffers: ONE small copy
        total, image = out[:4], out[4:].view(-1, 3)
        ctx.mark_non_differentiable(image, total)
        return total[0], total[1], image, total

    @staticmethod
    def backward(ctx, g_cls, g_reg, _g_image, _g_total):
        g = ctx.graph
        if g.steps_done != ctx.step_id or g.grads_taken:
            raise RuntimeError("RetinaNetLosses(graph=True): the gradient buffers of this forward are gone (a newer "
                               "forward replayed the graph, or backward already ran); use graph=False for retain_graph / "
                               "repeated backward")
        g.grads_taken = True
        outs = []
        for i, (buf, go) in enumerate(((g.grad_cls_preds, g_cls), (g.grad_bbox_preds, g_reg))):
            if go is None or not ctx.needs_input_grad[i]:
                outs.append(None)
                continue
            _scale_in_place(buf, go)
            outs.append(buf if ctx.in_dtypes[i] == torch.float32 else buf.to(ctx.in_dtypes[i]))
        return outs[0], outs[1], None, None


class RetinaNetLosses(nn.Module):
    """Reference: retinanet/losses.py:11-145 (hyper-parameters read from config at construction).

    ``graph=True`` (opt-in, same results bit for bit): ``forward`` replays a CUDA graph of the training half captured
    for the current input ADDRESSES and shapes (the caching allocator hands a model the same blocks step after step;
    a new address or shape captures another graph, a few are kept).  Host cost per call drops from ~0.4 ms to target
    packing + one graph launch.  Restrictions of this mode: fp32 contiguous head outputs, shared anchors, one backward
    per forward, and the gradients handed to autograd are the graph's static buffers (overwritten by the next
    forward — fine for a training step, not for code that keeps ``cls_preds.grad`` across steps)."""

    def __init__(self, num_classes: int, graph: bool = False) -> None:
        super().__init__()
        self.n_c = num_classes
        self.alpha = FOCAL_LOSS_ALPHA
        self.gamma = FOCAL_LOSS_GAMMA
        self.beta = SMOOTH_L1_LOSS_BETA
        self.graph = bool(graph)
        self.graph_max_targets = 8192
        self._graphs: Dict[tuple, object] = {}

    def _forward_graph(self, targets, clas_preds: Tensor, bbox_preds: Tensor, an: Tensor) -> Optional[Dict[str, Tensor]]:
        """Graph replay, or None when this call does not qualify (then the eager path runs)."""
        x, b = clas_preds, bbox_preds
        if not (x.is_cuda and x.dtype == torch.float32 and b.dtype == torch.float32 and x.is_contiguous()
                and b.is_contiguous() and len(targets) == x.shape[0] and x.shape[0] > 0):
            return None
        if sum(int(t["boxes"].shape[0]) for t in targets) > self.graph_max_targets:
            return None
        key = (x.data_ptr(), b.data_ptr(), tuple(x.shape), an.data_ptr(), x.device.index,
               self.alpha, self.gamma, self.beta, torch.cuda.current_stream(x.device).cuda_stream)
        g = self._graphs.get(key)
        if g is None:
            from .graphs import HotPathGraph
            if len(self._graphs) >= 4:
                self._graphs.pop(next(iter(self._graphs)))
            g = HotPathGraph(self.n_c, x.detach(), b.detach(), an, train=True, detect=False,
                             max_targets=self.graph_max_targets, alpha=self.alpha, gamma=self.gamma, beta=self.beta)
            g.release_inputs()                      # keep addresses, not tensors: the allocator must be free to recycle them
            self._graphs[key] = g
        c, r, image, _ = _GraphLoss.apply(x, b, g, targets)
        self.last_per_image = image
        return {"classification_loss": c, "regression_loss": r}

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
        if "cls_levels" in head_outputs:        # raw per-level conv outputs: no permute/cat pass
            return self.forward_levels(targets, head_outputs["cls_levels"], head_outputs["bbox_levels"], anchors)
        clas_preds, bbox_preds = head_outputs["cls_preds"], head_outputs["bbox_preds"]
        if clas_preds.shape[-1] != self.n_c:
            raise ValueError(f"cls_preds has {clas_preds.shape[-1]} classes, expected {self.n_c}")
        an, stride = _shared_anchors(anchors)
        if self.graph and stride == 0:
            out = self._forward_graph(targets, clas_preds, bbox_preds, an)
            if out is not None:
                return out
        packed = PackedTargets([t["boxes"] for t in targets], [t["labels"] for t in targets], clas_preds.device)
        c, r, image, _ = _FusedRetinaNetLoss.apply(clas_preds, bbox_preds, an, stride, packed, self._hp(len(targets)))
        self.last_per_image = image   # [N,3]: cls_i, reg_i, F_i (device tensor, no sync)
        return {"classification_loss": c, "regression_loss": r}

    def forward_levels(self, targets: List[Dict[str, Tensor]], cls_levels: Sequence[Tensor], bbox_levels: Sequence[Tensor],
                       anchors: List[Tensor], hp_extra: Optional[dict] = None) -> Dict[str, Tensor]:
        """Same loss as :meth:`forward`, computed directly on the head's per-level conv outputs
        ``cls_levels[l] = [N, na*C, H_l, W_l]`` / ``bbox_levels[l] = [N, na*4, H_l, W_l]`` (what
        ``RetinaNetClassSubnet`` / ``RetinaNetBoxSubnet`` hold before ``view/permute/contiguous/cat``,
        layers.py:189-195, 253-259).  Gradients come back in the same layout."""
        an, stride = _shared_anchors(anchors)
        packed = PackedTargets([t["boxes"] for t in targets], [t["labels"] for t in targets], cls_levels[0].device)
        hp = self._hp(len(targets))
        if hp_extra:
            hp.update(hp_extra)
        L = len(cls_levels)
        c, r, image, total = _FusedRetinaNetLossLevels.apply(an, stride, packed, hp, self.n_c, L, *cls_levels, *bbox_levels)
        self.last_per_image = image
        self.last_stats = total
        return {"classification_loss": c, "regression_loss": r}
