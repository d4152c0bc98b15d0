"""retinanet_b200 — B200-native (sm_100a) dense per-anchor path of RetinaNet.

Drop-in replacements for the hot-path pieces of benihime91/pytorch_retinanet, same names and
signatures: ``AnchorGenerator`` (retinanet/anchors.py), ``matcher`` / ``bbox_2_activ`` /
``activ_2_bbox`` (retinanet/box_utils.py), ``RetinaNetLosses`` (retinanet/losses.py) and
``process_detections`` (retinanet/models.py:160-243).  Every stage is a hand-written CUDA kernel
behind the C ABI of ``include/retinanet_b200.h``; there is no CPU fallback.
"""
from .anchors import AnchorGenerator, BufferList
from .box_utils import activ_2_bbox, bbox_2_activ, convert_x1y1x2y2, convert_xywh, matcher
from .detections import (PendingDetections, postprocess_batch, postprocess_batch_async, process_detections,
                         process_detections_async)
from .graphs import HotPathGraph, HotPathPipeline
from .integration import patch_retinanet
from .losses import RetinaNetLosses

__all__ = ["AnchorGenerator", "BufferList", "matcher", "bbox_2_activ", "activ_2_bbox", "convert_xywh",
           "convert_x1y1x2y2", "RetinaNetLosses", "process_detections", "process_detections_async", "postprocess_batch", "postprocess_batch_async",
           "PendingDetections", "patch_retinanet", "HotPathGraph", "HotPathPipeline"]
# image-sharded multi-GPU use: pytorch_retinanet_b200.distributed (ShardedRetinaNetLosses, PeerExchange) — imported on
# demand, it pulls in torch.distributed
