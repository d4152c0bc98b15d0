"""Image-sharded multi-GPU loss: one process per GPU, each rank owns a contiguous
slice of the batch; images are independent in the reference (python loop, per-image normaliser,
retinanet/losses.py:126-140), so the only exchange is ONE all-reduce (NCCL over NVLink) of four
floats per step: [sum_i cls_i/max(1,F_i), sum_i reg_i/max(1,F_i), sum_i F_i, N_local].
Inference post-processing needs no communication at all.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist
from torch import Tensor

from .box_utils import PackedTargets
from .losses import RetinaNetLosses, _FusedRetinaNetLoss, _shared_anchors


def shard_range(num_images: int, rank: int, world: int):
    """Rank r owns images [r*N/W, (r+1)*N/W) (contiguous, sizes differ by at most one)."""
    return (num_images * rank) // world, (num_images * (rank + 1)) // world


class ShardedRetinaNetLosses(RetinaNetLosses):
    """``forward`` takes this rank's shard and returns the loss of the GLOBAL batch (identical on
    every rank and equal to the single-process reference on the full batch up to fp32 summation
    order).  Gradients w.r.t. the local shard are already scaled by 1/N_global."""

    def __init__(self, num_classes: int, global_batch: Optional[int] = None, group=None) -> None:
        super().__init__(num_classes)
        self.global_batch = global_batch
        self.group = group
        self.last_stats: Optional[Tensor] = None

    def forward(self, targets: List[Dict[str, Tensor]], head_outputs: Dict[str, Tensor],
                anchors: List[Tensor]) -> Dict[str, Tensor]:
        cls, box = head_outputs["cls_preds"], head_outputs["bbox_preds"]
        n_local = len(targets)
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        n_global = self.global_batch
        if n_global is None:
            if world > 1:
                t = torch.tensor([float(n_local)], device=cls.device)
                dist.all_reduce(t, group=self.group)
                n_global = int(t.item())
            else:
                n_global = n_local
        an, stride = _shared_anchors(anchors)
        packed = PackedTargets([t["boxes"] for t in targets], [t["labels"] for t in targets], cls.device)
        hp = self._hp(n_global)
        if world > 1:
            hp["all_reduce_group"] = self.group                # the all-reduce happens inside the autograd function
        c, r, image, total = _FusedRetinaNetLoss.apply(cls, box, an, stride, packed, hp)
        self.last_per_image = image
        self.last_stats = total                                # [cls, reg, sum F, N] of the GLOBAL batch (device tensor)
        return {"classification_loss": c, "regression_loss": r}
