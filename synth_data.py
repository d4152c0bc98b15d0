"""Seeded synthetic inputs for the five BASELINE.json configs.

Self-contained data generation shared by ``bench.py`` (both arms) and ``tests/``;
it is not part of the product path and does not import ``oracle/``.  Everything is
generated on the CPU from ``torch.Generator().manual_seed(1000*config_id + image_index)``
so the data is identical for every world size / sharding and for CPU and GPU runs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

STRIDES = [8, 16, 32, 64, 128]


@dataclass(frozen=True)
class PathConfig:
    config_id: int
    name: str
    im_hw: Tuple[int, int]          # resized, unpadded image size handed to clip (h, w)
    padded_hw: Tuple[int, int]      # size after GeneralizedRCNNTransform padding (multiple of 32)
    num_classes: int
    batch: int
    gt_range: Tuple[int, int]       # inclusive range of GT boxes per image
    clustered: bool


CONFIGS: Dict[int, PathConfig] = {
    1: PathConfig(1, "voc512_c20_g10_n1", (512, 512), (512, 512), 20, 1, (10, 10), False),
    2: PathConfig(2, "coco800x1333_c80_g<=100_n16_train", (800, 1333), (800, 1344), 80, 16, (1, 100), True),
    3: PathConfig(3, "coco800x1333_c80_g<=100_n128_sharded", (800, 1333), (800, 1344), 80, 128, (1, 100), True),
    4: PathConfig(4, "coco800x1333_c80_n256_infer", (800, 1333), (800, 1344), 80, 256, (1, 100), True),
    5: PathConfig(5, "crowd1024_c80_g500_n64", (1024, 1024), (1024, 1024), 80, 64, (500, 500), True),
}


def grid_sizes(padded_hw: Tuple[int, int]) -> List[Tuple[int, int]]:
    h, w = padded_hw
    return [(-(-h // s), -(-w // s)) for s in STRIDES]


def num_anchors(padded_hw: Tuple[int, int]) -> int:
    return sum(9 * h * w for h, w in grid_sizes(padded_hw))


def default_anchors(padded_hw: Tuple[int, int]) -> torch.Tensor:
    """Default P3–P7 anchor set on the CPU (sizes 32..512 × 2^{0,1/3,2/3}, ratios .5/1/2, offset 0)."""
    out = []
    for (H, W), stride, base in zip(grid_sizes(padded_hw), STRIDES, [32, 64, 128, 256, 512]):
        rows = []
        for s in (base, base * 2 ** (1 / 3), base * 2 ** (2 / 3)):
            for r in (0.5, 1.0, 2.0):
                w = math.sqrt(s ** 2.0 / r)
                h = r * w
                rows.append([-w / 2.0, -h / 2.0, w / 2.0, h / 2.0])
        cell = torch.tensor(rows).float()
        sx = torch.arange(0, W * stride, step=stride, dtype=torch.float32)
        sy = torch.arange(0, H * stride, step=stride, dtype=torch.float32)
        gy, gx = torch.meshgrid(sy, sx, indexing="ij")
        sh = torch.stack((gx.reshape(-1), gy.reshape(-1), gx.reshape(-1), gy.reshape(-1)), 1)
        out.append((sh[:, None, :] + cell[None]).reshape(-1, 4))
    return torch.cat(out)


def nac_to_levels(x: torch.Tensor, padded_hw: Tuple[int, int], na: int = 9) -> List[torch.Tensor]:
    """[N, A, K] (the reference head's output layout) -> the per-level conv outputs [N, na*K, H_l, W_l]
    it was permuted from (inverse of retinanet/layers.py:189-195); used to feed the row-N1 entry points."""
    N, _, K = x.shape
    out, off = [], 0
    for h, w in grid_sizes(padded_hw):
        n = h * w * na
        out.append(x[:, off:off + n].reshape(N, h, w, na, K).permute(0, 3, 4, 1, 2).reshape(N, na * K, h, w).contiguous())
        off += n
    return out


def _gt_boxes(g: torch.Generator, n: int, hw: Tuple[int, int]) -> torch.Tensor:
    H, W = hw
    boxes = torch.empty((n, 4), dtype=torch.float32)
    todo = torch.arange(n)
    while todo.numel():
        k = todo.numel()
        cx = torch.rand(k, generator=g) * W
        cy = torch.rand(k, generator=g) * H
        scale = 2.0 ** (4.0 + 5.0 * torch.rand(k, generator=g))
        r = 2.0 ** (-1.58 + 3.16 * torch.rand(k, generator=g))
        w, h = scale / r.sqrt(), scale * r.sqrt()
        b = torch.stack([(cx - w / 2).clamp(0, W), (cy - h / 2).clamp(0, H),
                         (cx + w / 2).clamp(0, W), (cy + h / 2).clamp(0, H)], 1)
        ok = ((b[:, 2] - b[:, 0]) >= 2) & ((b[:, 3] - b[:, 1]) >= 2)
        boxes[todo[ok]] = b[ok]
        todo = todo[~ok]
    return boxes


def _best_gt(anchors: torch.Tensor, gt: torch.Tensor, chunk: int = 32768):
    """max-IoU GT per anchor (data generation only)."""
    vals = torch.empty(anchors.shape[0])
    idxs = torch.empty(anchors.shape[0], dtype=torch.int64)
    ag = (gt[:, 2] - gt[:, 0]) * (gt[:, 3] - gt[:, 1])
    for s in range(0, anchors.shape[0], chunk):
        a = anchors[s:s + chunk]
        aa = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
        wh = (torch.min(gt[:, None, 2:], a[None, :, 2:]) - torch.max(gt[:, None, :2], a[None, :, :2])).clamp(min=0)
        inter = wh[..., 0] * wh[..., 1]
        v, i = (inter / (ag[:, None] + aa[None] - inter)).max(0)
        vals[s:s + chunk], idxs[s:s + chunk] = v, i
    return vals, idxs


def make_image(cfg: PathConfig, image_index: int, anchors: torch.Tensor, clustered: Optional[bool] = None):
    """One image's inputs: (logits [A,C], bbox [A,4], gt [G,4], labels [G] int64 1-based)."""
    clustered = cfg.clustered if clustered is None else clustered
    g = torch.Generator().manual_seed(1000 * cfg.config_id + image_index)
    A, C = anchors.shape[0], cfg.num_classes
    lo, hi = cfg.gt_range
    G = lo if lo == hi else int(torch.randint(lo, hi + 1, (1,), generator=g))
    gt = _gt_boxes(g, G, cfg.im_hw)
    labels = torch.randint(1, C + 1, (G,), generator=g, dtype=torch.int64)
    logits = torch.randn((A, C), generator=g).mul_(1.3).add_(-7.0)
    bbox = torch.randn((A, 4), generator=g).mul_(0.1)
    if clustered and G > 0:
        vals, idxs = _best_gt(anchors, gt)
        fg = torch.nonzero(vals > 0.5).squeeze(1)
        if fg.numel():
            gi = idxs[fg]
            logits[fg, labels[gi] - 1] += torch.randn(fg.numel(), generator=g) * 1.5 + 5.0
            a, b = anchors[fg], gt[gi]
            aw, ah = a[:, 2] - a[:, 0], a[:, 3] - a[:, 1]
            ax, ay = (a[:, 0] + a[:, 2]) / 2, (a[:, 1] + a[:, 3]) / 2
            bw, bh = b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]
            bx, by = (b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2
            tgt = torch.stack([(bx - ax) / aw, (by - ay) / ah, torch.log(bw / aw + 1e-8), torch.log(bh / ah + 1e-8)], 1)
            bbox[fg] = tgt + torch.randn((fg.numel(), 4), generator=g) * 0.05
    return logits, bbox, gt, labels


def make_batch(cfg: PathConfig, first_image: int = 0, count: Optional[int] = None,
               clustered: Optional[bool] = None, pin: bool = False):
    """Inputs for images [first_image, first_image+count) of ``cfg`` (all CPU tensors).

    Returns dict with ``cls_preds [n,A,C]``, ``bbox_preds [n,A,4]``, ``anchors [A,4]``,
    ``targets`` (list of {"boxes","labels"}), ``im_szs`` (list of (h,w)).
    """
    count = cfg.batch if count is None else count
    anchors = default_anchors(cfg.padded_hw)
    A, C = anchors.shape[0], cfg.num_classes
    cls = torch.empty((count, A, C), dtype=torch.float32, pin_memory=pin)
    box = torch.empty((count, A, 4), dtype=torch.float32, pin_memory=pin)
    targets = []
    for i in range(count):
        lg, bb, gt, lab = make_image(cfg, first_image + i, anchors, clustered)
        cls[i], box[i] = lg, bb
        targets.append({"boxes": gt, "labels": lab})
    return {"cls_preds": cls, "bbox_preds": box, "anchors": anchors, "targets": targets,
            "im_szs": [cfg.im_hw] * count, "config": cfg}


# ---- the same distributions generated ON the device (bench legs whose batches are too large to draw on the host
# within the bench's time budget: 256 images of config 4, 128 of config 3, 64 of config 5).  Different random streams
# than make_batch (CUDA generator), same per-image seeding rule and shapes; parity tests always use make_batch. ----
def make_batch_device(cfg: PathConfig, first_image: int, count: int, device, clustered: Optional[bool] = None,
                      want_bbox: bool = True):
    """Inputs for images [first_image, first_image+count) drawn with a CUDA generator seeded per image
    (1000*config_id + image_index).  Returns the same dict as :func:`make_batch` with CUDA tensors."""
    clustered = cfg.clustered if clustered is None else clustered
    device = torch.device(device)
    anchors = default_anchors(cfg.padded_hw).to(device)
    A, C = anchors.shape[0], cfg.num_classes
    cls = torch.empty((count, A, C), dtype=torch.float32, device=device)
    box = torch.empty((count, A, 4), dtype=torch.float32, device=device) if want_bbox else None
    H, W = cfg.im_hw
    targets = []
    g = torch.Generator(device=device)
    for i in range(count):
        g.manual_seed(1000 * cfg.config_id + first_image + i)
        lo, hi = cfg.gt_range
        G = lo if lo == hi else lo + (first_image + i) * 7919 % (hi - lo + 1)
        k = G * 2 + 8                                            # oversample, keep the first G valid boxes
        cx = torch.rand(k, generator=g, device=device) * W
        cy = torch.rand(k, generator=g, device=device) * H
        scale = 2.0 ** (4.0 + 5.0 * torch.rand(k, generator=g, device=device))
        r = 2.0 ** (-1.58 + 3.16 * torch.rand(k, generator=g, device=device))
        w, h = scale / r.sqrt(), scale * r.sqrt()
        b = torch.stack([(cx - w / 2).clamp(0, W), (cy - h / 2).clamp(0, H), (cx + w / 2).clamp(0, W), (cy + h / 2).clamp(0, H)], 1)
        ok = ((b[:, 2] - b[:, 0]) >= 2) & ((b[:, 3] - b[:, 1]) >= 2)
        gt = b[ok][:G].contiguous()
        G = gt.shape[0]
        labels = torch.randint(1, C + 1, (G,), generator=g, device=device, dtype=torch.int64)
        cls[i].normal_(-7.0, 1.3, generator=g)
        if want_bbox:
            box[i].normal_(0.0, 0.1, generator=g)
        if clustered and G > 0:
            vals = torch.empty(A, device=device)
            idxs = torch.empty(A, dtype=torch.int64, device=device)
            ag = (gt[:, 2] - gt[:, 0]) * (gt[:, 3] - gt[:, 1])
            for s0 in range(0, A, 65536):
                a = anchors[s0:s0 + 65536]
                aa = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
                wh = (torch.min(gt[:, None, 2:], a[None, :, 2:]) - torch.max(gt[:, None, :2], a[None, :, :2])).clamp(min=0)
                inter = wh[..., 0] * wh[..., 1]
                v, ix = (inter / (ag[:, None] + aa[None] - inter)).max(0)
                vals[s0:s0 + 65536], idxs[s0:s0 + 65536] = v, ix
            fg = torch.nonzero(vals > 0.5).squeeze(1)
            if fg.numel():
                gi = idxs[fg]
                cls[i][fg, labels[gi] - 1] += torch.randn(fg.numel(), generator=g, device=device) * 1.5 + 5.0
                if want_bbox:
                    a, bb = anchors[fg], gt[gi]
                    aw, ah = a[:, 2] - a[:, 0], a[:, 3] - a[:, 1]
                    ax, ay = (a[:, 0] + a[:, 2]) / 2, (a[:, 1] + a[:, 3]) / 2
                    bw, bh = bb[:, 2] - bb[:, 0], bb[:, 3] - bb[:, 1]
                    bx, by = (bb[:, 0] + bb[:, 2]) / 2, (bb[:, 1] + bb[:, 3]) / 2
                    tgt = torch.stack([(bx - ax) / aw, (by - ay) / ah, torch.log(bw / aw + 1e-8), torch.log(bh / ah + 1e-8)], 1)
                    box[i][fg] = tgt + torch.randn((fg.numel(), 4), generator=g, device=device) * 0.05
        targets.append({"boxes": gt, "labels": labels})
    return {"cls_preds": cls, "bbox_preds": box, "anchors": anchors, "targets": targets,
            "im_szs": [cfg.im_hw] * count, "config": cfg}
