"""Drop-in ``AnchorGenerator`` (reference: retinanet/anchors.py) backed by one CUDA launch.

Same constructor, properties, buffer names (``cell_anchors.0..4`` — state_dict compatible) and
return types.  Differences, all invisible to the reference's call sites (models.py:266,284):
the grid for a given set of feature-map sizes is generated ONCE per forward (the reference
regenerates the identical tensor for every image, anchors.py:223-226) and the same tensor object is
returned for every image; results are cached per (grid sizes, device).
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, List, Tuple

import torch
from torch import Tensor, nn

from . import _native
from .config import ANCHOR_ASPECT_RATIOS, ANCHOR_OFFSET, ANCHOR_SIZES, ANCHOR_STRIDES
from .utilities import ifnone


class BufferList(nn.Module):
    """Buffers registered under the names "0", "1", ... (reference: anchors.py:13-27)."""

    def __init__(self, buffers):
        super().__init__()
        for i, b in enumerate(buffers):
            self.register_buffer(str(i), b)

    def __len__(self):
        return len(self._buffers)

    def __iter__(self):
        return iter(self._buffers.values())


def _broadcast_params(params, num_features, name) -> List[List[float]]:
    """Reference: anchors.py:30-52 (same assertions and messages' meaning)."""
    assert isinstance(params, (list, tuple)), f"{name} in anchor generator has to be a list! Got {params}."
    assert len(params), f"{name} in anchor generator cannot be empty!"
    if not isinstance(params[0], (list, tuple)):
        return [params] * num_features
    if len(params) == 1:
        return list(params) * num_features
    assert len(params) == num_features, (
        f"Got {name} of length {len(params)} in anchor generator, "
        f"but the number of input features is {num_features}!"
    )
    return params


class AnchorGenerator(nn.Module):
    """Generates anchors for a set of feature maps (reference: anchors.py:55-228)."""

    def __init__(self, sizes=None, aspect_ratios=None, strides=None, offset=None) -> None:
        super().__init__()
        strides = ifnone(strides, ANCHOR_STRIDES)
        sizes = ifnone(sizes, ANCHOR_SIZES)
        aspect_ratios = ifnone(aspect_ratios, ANCHOR_ASPECT_RATIOS)
        offset = ifnone(offset, ANCHOR_OFFSET)
        self.strides = strides
        self.num_features = len(strides)
        self.sizes = _broadcast_params(sizes, self.num_features, "sizes")
        self.aspect_ratios = _broadcast_params(aspect_ratios, self.num_features, "aspect_ratios")
        self.offset = offset
        self.cell_anchors = self._calculate_cell_anchors(self.sizes, self.aspect_ratios)
        self._cache: Dict[Tuple, Tuple] = {}     # key -> (anchors, ready event, producing stream)
        self.last_level_offsets: List[int] = []   # anchor offsets of the pyramid levels of the last grid

    def _calculate_cell_anchors(self, sizes, ratios):
        return self._calculate_anchors(sizes, ratios)

    def _calculate_anchors(self, sizes, aspect_ratios) -> BufferList:
        return BufferList([self.generate_cell_anchors(s, a).float() for s, a in zip(sizes, aspect_ratios)])

    @staticmethod
    def generate_cell_anchors(sizes, aspect_ratios) -> Tensor:
        """Cell anchors centred at (0,0), size-major / ratio-minor, computed in Python double on the
        host exactly as the reference does (anchors.py:110-135): 45 numbers, not a kernel's job."""
        rows = []
        for size in sizes:
            area = size ** 2.0
            for r in aspect_ratios:
                w = math.sqrt(area / r)
                h = r * w
                rows.append([-w / 2.0, -h / 2.0, w / 2.0, h / 2.0])
        return torch.tensor(rows)

    @property
    def num_cell_anchors(self):
        return self.num_anchors

    @property
    def num_anchors(self) -> List[int]:
        return [len(c) for c in self.cell_anchors]

    def _apply(self, fn, *a, **k):  # .to()/.cuda() invalidates cached grids
        self._cache = {}
        return super()._apply(fn, *a, **k)

    def _all_levels(self, grid_sizes, device: torch.device) -> Tensor:
        """[A,4] anchors of all levels (concatenated) from one ``rn_anchor_grid`` launch."""
        device = torch.device(device)
        if device.type != "cuda":
            raise _native.NativeError("retinanet_b200.AnchorGenerator: feature maps must live on a CUDA device")
        grid_sizes = [(int(h), int(w)) for h, w in grid_sizes]
        offs = [0]
        for (h, w), n in zip(grid_sizes, self.num_anchors):
            offs.append(offs[-1] + h * w * n)
        self.last_level_offsets = offs
        cells = [b for b in self.cell_anchors]
        # the key covers everything the grid depends on: a load_state_dict into the cell_anchors buffers bumps their
        # version counters, strides / offset are compared by value — a stale grid is never served
        key = (tuple(grid_sizes), device.index if device.index is not None else torch.cuda.current_device(),
               tuple((c.data_ptr(), c._version, tuple(c.shape)) for c in cells),
               tuple(int(s) for s in self.strides), float(self.offset))
        stream = torch.cuda.current_stream(device)
        hit = self._cache.get(key)
        if hit is not None:
            anchors, ready, made_on = hit
            if made_on != stream.cuda_stream:                  # first use on another stream: order after the producer
                stream.wait_event(ready)
            return anchors
        if len(self._cache) > 16:
            self._cache.clear()
        assert len(grid_sizes) == self.num_features, "one feature map per stride expected"
        cells_dev = torch.cat([c.to(device=device, dtype=torch.float32) for c in cells]).contiguous()
        desc = []
        total = 0
        for (h, w), st, c in zip(grid_sizes, self.strides, cells):
            desc += [h, w, int(st), c.shape[0]]
            total += h * w * c.shape[0]
        out = torch.empty((total, 4), dtype=torch.float32, device=device)
        lib = _native.load()
        with _native.on_device(device):
            rc = lib.rn_anchor_grid(_native.ptr(cells_dev), (ctypes.c_int32 * len(desc))(*desc), len(grid_sizes),
                                    float(self.offset), _native.ptr(out), total, _native.stream_ptr(device))
        _native.check(rc, "rn_anchor_grid")
        ready = torch.cuda.Event()
        ready.record(stream)
        self._cache[key] = (out, ready, stream.cuda_stream)
        return out

    def grid_anchors(self, grid_sizes, device) -> List[Tensor]:
        """One tensor [(H_l*W_l*na_l), 4] per level (views of one buffer)."""
        allv = self._all_levels(grid_sizes, device)
        counts = [int(h) * int(w) * n for (h, w), n in zip(grid_sizes, self.num_anchors)]
        return list(torch.split(allv, counts))

    def forward(self, images, feature_maps: List[Tensor]) -> List[Tensor]:
        """``images``: ImageList (only ``len(images.image_sizes)`` is used, as in anchors.py:223)."""
        grid_sizes = [fm.shape[-2:] for fm in feature_maps]
        anchors = self._all_levels(grid_sizes, feature_maps[0].device)
        return [anchors for _ in images.image_sizes]
