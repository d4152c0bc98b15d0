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


def _head_forward_levels(head, feature_maps):
    """Replacement for ``RetinaNetHead.forward`` (layers.py:110-115) that stops BEFORE the subnets'
    ``view -> permute(0,3,4,1,2) -> contiguous -> cat`` (layers.py:189-195, 253-259): it returns the raw
    per-level conv outputs, which our loss / post-processing kernels index directly (row N1)."""
    cls_sub, box_sub = head.classification_head, head.regression_head
    cls_levels = [cls_sub.class_subnet_output(cls_sub.class_subnet(f)) for f in feature_maps]
    box_levels = [box_sub.box_subnet_output(box_sub.box_subnet(f)) for f in feature_maps]
    return {"cls_levels": cls_levels, "bbox_levels": box_levels}


def patch_retinanet(model, pre_nms_topk=None, fuse_head_layout=False):
    """In-place swap; returns ``model``.  Existing ``cell_anchors`` buffers are carried over so that
    state_dict keys (``anchor_generator.cell_anchors.{i}``) and values are unchanged.
    ``fuse_head_layout=True`` additionally makes the head hand over its raw per-level conv outputs
    (no permute/contiguous/cat pass over the logits)."""
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
    if fuse_head_layout:
        model.retinanet_head.forward = types.MethodType(_head_forward_levels, model.retinanet_head)
    return model
