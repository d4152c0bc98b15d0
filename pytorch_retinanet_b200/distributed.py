"""Image-sharded multi-GPU loss: one process per GPU, each rank owns a contiguous
slice of the batch; images are independent in the reference (python loop, per-image normaliser,
retinanet/losses.py:126-140), so the only exchange is the sum over the ranks of four floats per step:
[sum_i cls_i/max(1,F_i), sum_i reg_i/max(1,F_i), sum_i F_i, N_local].
Inference post-processing needs no communication at all.

The exchange is folded into the loss's final reduction kernel (:class:`PeerExchange`): the block that finishes the
local sum stores its vector into every peer's receive slots over NVLink (CUDA IPC peer-mapped memory) and adds up
what the peers stored for it — no separate collective launch, nothing on the host, CUDA-graph capturable.
``torch.distributed`` only ships the 64-byte IPC handles once, at set-up.  ``exchange="nccl"`` keeps the one
``all_reduce`` per step instead (also what the gloo CPU tests use).
"""
from __future__ import annotations

import ctypes
import os
import socket
import warnings
from typing import Dict, List, Optional

import torch
import torch.distributed as dist
from torch import Tensor

from . import _native
from .box_utils import PackedTargets
from .losses import RetinaNetLosses, _FusedRetinaNetLoss, _shared_anchors


def shard_range(num_images: int, rank: int, world: int):
    """Rank r owns images [r*N/W, (r+1)*N/W) (contiguous, sizes differ by at most one)."""
    return (num_images * rank) // world, (num_images * (rank + 1)) // world


class PeerExchange:
    """Receive buffers of the in-kernel loss exchange, mapped across the ranks of ``group`` (one NVLink/NVSwitch box).

    Collective constructor: every rank allocates its 2 KB buffer (``rn_comm_alloc``), the 64-byte CUDA IPC handles go
    round with ``all_gather_object`` and every rank maps every peer's buffer.  ``ref`` is the ``rn_exchange_t *`` the
    loss entry points take.  All ranks must then issue the same sequence of exchanging calls."""

    def __init__(self, group=None, device: Optional[torch.device] = None):
        if not dist.is_initialized():
            raise RuntimeError("PeerExchange needs an initialised torch.distributed process group")
        lib = _native.load()
        self.lib, self.group = lib, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > _native.RN_MAX_PEERS:
            raise _native.NativeError(f"PeerExchange: {self.world} ranks exceed RN_MAX_PEERS={_native.RN_MAX_PEERS}")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._local = ctypes.c_void_p()
        self._mapped: List[int] = []
        handle = (ctypes.c_ubyte * 64)()
        err = None
        try:
            with torch.cuda.device(self.device):
                _native.check(lib.rn_comm_alloc(ctypes.byref(self._local)), "rn_comm_alloc")
                _native.check(lib.rn_comm_export(self._local, handle), "rn_comm_export")
        except _native.NativeError as e:
            err = str(e)
        mine = (bytes(handle), socket.gethostname(), err)
        everyone: List = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        fail = next((f"rank {r}: {e[2]}" for r, e in enumerate(everyone) if e[2]), None)
        if fail is None and len({e[1] for e in everyone}) != 1:
            fail = "ranks live on different hosts (peer-mapped memory needs one NVLink box)"
        peers = [None] * self.world
        if fail is None:
            try:
                with torch.cuda.device(self.device):
                    for r, (h, _, _) in enumerate(everyone):
                        if r == self.rank:
                            peers[r] = self._local.value
                            continue
                        p = ctypes.c_void_p()
                        _native.check(lib.rn_comm_import(h, ctypes.byref(p)), f"rn_comm_import(rank {r})")
                        self._mapped.append(p.value)
                        peers[r] = p.value
            except _native.NativeError as e:
                fail = f"rank {self.rank}: {e}"
        flags: List = [None] * self.world                      # second round: did EVERY rank map EVERY buffer?
        dist.all_gather_object(flags, fail, group=group)
        fail = next((f for f in flags if f), None)
        if fail is not None:
            self.close(collective=False)
            raise _native.NativeError(f"PeerExchange set-up failed ({fail}); use exchange='nccl'")
        self.struct = _native.RnExchange()
        for r, p in enumerate(peers):
            self.struct.peers[r] = p
        self.struct.rank, self.struct.world = self.rank, self.world
        self.ref = ctypes.byref(self.struct)

    def exchange(self, total: Tensor) -> Tensor:
        """total[4] <- sum over the ranks, in place (one tiny launch); what the loss kernels do internally."""
        with _native.on_device(total.device):
            rc = self.lib.rn_exchange_total(_native.ptr(total, torch.float32, "total"), self.ref,
                                            _native.stream_ptr(total.device))
        _native.check(rc, "rn_exchange_total")
        return total

    def error(self) -> bool:
        """True if a peer failed to arrive within the kernel's time-out in some earlier step (synchronises)."""
        v = ctypes.c_int32(0)
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            _native.check(self.lib.rn_comm_error(self._local, ctypes.byref(v)), "rn_comm_error")
        return bool(v.value)

    def close(self, collective: bool = True) -> None:
        """Unmaps the peers' buffers and frees the local one.  Collective by default: nobody frees a buffer a peer may
        still store into."""
        if self._local is None:
            return
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            if collective and dist.is_initialized():
                dist.barrier(group=self.group)
            for p in self._mapped:
                self.lib.rn_comm_unmap(p)
            if self._local.value:
                self.lib.rn_comm_free(self._local)
        self._mapped, self._local = [], None


_EXCHANGES: Dict = {}


def get_exchange(group=None, mode: str = "peer"):
    """The process-wide :class:`PeerExchange` of ``group`` (created collectively on first use), or None for
    ``mode="nccl"`` / a single rank.  ``mode="auto"`` falls back to NCCL — on every rank alike, with a warning —
    when the peer mapping cannot be set up."""
    if mode not in ("peer", "nccl", "auto"):
        raise ValueError(f"exchange must be 'peer', 'nccl' or 'auto' (got {mode!r})")
    if mode == "nccl" or not dist.is_initialized() or dist.get_world_size(group) <= 1:
        return None
    key = (id(group) if group is not None else None, torch.cuda.current_device())
    if key not in _EXCHANGES:
        try:
            _EXCHANGES[key] = PeerExchange(group)
        except _native.NativeError as e:            # raised on every rank alike (the constructor agrees collectively)
            if mode == "peer":
                raise
            warnings.warn(f"retinanet_b200: {e}; the loss vector goes through one NCCL all_reduce per step instead")
            _EXCHANGES[key] = None
    return _EXCHANGES[key]


def close_exchanges() -> None:
    """Collective: releases every :class:`PeerExchange` made by :func:`get_exchange` (call before
    ``destroy_process_group``)."""
    for x in list(_EXCHANGES.values()):
        if x is not None:
            x.close()
    _EXCHANGES.clear()


class ShardedRetinaNetLosses(RetinaNetLosses):
    """``forward`` takes this rank's shard and returns the loss of the GLOBAL batch (identical on
    every rank and equal to the single-process reference on the full batch up to fp32 summation
    order).

    Gradients w.r.t. the local shard are scaled by 1/N_global, i.e. they are this rank's TERM of the global
    gradient: parameter gradients must be SUMMED over the ranks.  ``grad_reduction="mean"`` is for wrappers that
    average instead (``DistributedDataParallel``, Lightning's ddp — what the reference's trainer would use): the local
    gradients are then scaled by world/N_global so that the average over the ranks is the global gradient; the
    returned loss values are the same in both modes.

    ``exchange``: "peer" (default on CUDA; the sum is done inside the loss kernel over peer-mapped memory), "nccl"
    (one ``all_reduce`` per step; the only choice for CPU/gloo tensors), "auto" (peer, NCCL if the mapping fails), or
    a :class:`PeerExchange`.  An empty shard (fewer images than ranks) still takes part in the exchange."""

    def __init__(self, num_classes: int, global_batch: Optional[int] = None, group=None, exchange="peer",
                 grad_reduction: str = "sum") -> None:
        super().__init__(num_classes)
        if grad_reduction not in ("sum", "mean"):
            raise ValueError("grad_reduction must be 'sum' or 'mean'")
        self.global_batch = global_batch
        self.group = group
        self.exchange = exchange
        self.grad_reduction = grad_reduction
        self.last_stats: Optional[Tensor] = None

    def _sharded_hp(self, n_local: int, device: torch.device) -> dict:
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        n_global = self.global_batch
        if n_global is None:
            if world > 1:
                t = torch.tensor([float(n_local)], device=device)
                dist.all_reduce(t, group=self.group)
                n_global = int(t.item())
            else:
                n_global = n_local
        hp = self._hp(n_global)
        if world > 1:
            if self.grad_reduction == "mean":
                hp["batch_div"] = float(n_global) / world     # local gradients x world; the loss value is put back below
                hp["total_scale"] = 1.0 / world
            x = self.exchange
            if not isinstance(x, PeerExchange):
                x = get_exchange(self.group, x) if device.type == "cuda" else None
            if x is not None:
                hp["exchange"] = x                              # summed inside the loss's final reduction kernel
            else:
                hp["all_reduce_group"] = self.group            # the all-reduce happens inside the autograd function
        return hp

    def forward(self, targets: List[Dict[str, Tensor]], head_outputs: Dict[str, Tensor],
                anchors: List[Tensor]) -> Dict[str, Tensor]:
        n_local = len(targets)
        if "cls_levels" in head_outputs:                       # raw per-level conv outputs (fuse_head_layout=True)
            cls_levels, box_levels = head_outputs["cls_levels"], head_outputs["bbox_levels"]
            hp = self._sharded_hp(n_local, cls_levels[0].device)
            return self.forward_levels(targets, cls_levels, box_levels, anchors,
                                       hp_extra={k: v for k, v in hp.items()
                                                 if k in ("batch_div", "total_scale", "exchange", "all_reduce_group")})
        cls, box = head_outputs["cls_preds"], head_outputs["bbox_preds"]
        hp = self._sharded_hp(n_local, cls.device)
        if n_local == 0:                                       # empty shard: no anchors / targets to look at
            an, stride = cls.new_zeros((max(cls.shape[1], 1), 4), dtype=torch.float32), 0
        else:
            an, stride = _shared_anchors(anchors)
        packed = PackedTargets([t["boxes"] for t in targets], [t["labels"] for t in targets], cls.device)
        c, r, image, total = _FusedRetinaNetLoss.apply(cls, box, an, stride, packed, hp)
        self.last_per_image = image
        self.last_stats = total                                # [cls, reg, sum F, N] of the GLOBAL batch (device tensor)
        return {"classification_loss": c, "regression_loss": r}
