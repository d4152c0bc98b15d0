"""Swap the hot path of an existing reference ``Retinanet`` instance for the CUDA one.

The reference reaches the path at four call sites (retinanet/models.py:266,270,284,287):
``self.anchor_generator(...)``, ``self.compute_loss(...)`` -> ``self.retinanet_head.losses(...)``
and ``self.process_detections(...)``.  ``patch_retinanet`` replaces exactly those three attributes
and nothing else (backbone, FPN, heads, transform stay the reference's).
"""
from __future__ import annotations

import types

from .anchors import AnchorGenerator
from .detections import process_detections
from .losses import RetinaNetLosses


def patch_retinanet(model, pre_nms_topk=None):
    """In-place swap; returns ``model``.  Existing ``cell_anchors`` buffers are carried over so that
    state_dict keys (``anchor_generator.cell_anchors.{i}``) and values are unchanged."""
    old = model.anchor_generator
    new = AnchorGenerator(sizes=getattr(old, "sizes", None), aspect_ratios=getattr(old, "aspect_ratios", None),
                          strides=getattr(old, "strides", None), offset=getattr(old, "offset", None))
    try:
        dev = next(iter(old.cell_anchors)).device
        new = new.to(dev)
    except StopIteration:
        pass
    model.anchor_generator = new
    model.retinanet_head.losses = RetinaNetLosses(model.num_classes)
    model.process_detections = types.MethodType(process_detections, model)
    if pre_nms_topk is not None:
        model.pre_nms_topk = pre_nms_topk
    return model
