"""The reference arm of ``bench.py`` and the real-model tests: imports the UNMODIFIED reference package and drives
its own public API for the path.  NOT part of the product (``pytorch_retinanet_b200`` never imports this).

Where the reference comes from, in order: ``$RN_REFERENCE_ROOT``, ``/root/reference`` (build container only),
``baseline/_ref`` (installed by ``baseline/install_reference.py``; git-ignored, travels to the GPU box).

The single import obstacle is ``retinanet/backbone.py:6`` (``from torchvision.models.utils import
load_state_dict_from_url`` — a module removed from current torchvision): a stub module is registered in
``sys.modules`` before the import; no reference file is touched.
"""
import os
import sys
import types
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = [os.environ.get("RN_REFERENCE_ROOT"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")]


def reference_root():
    for c in CANDIDATES:
        if c and os.path.isfile(os.path.join(c, "retinanet", "losses.py")):
            return c
    return None


def reference_available() -> bool:
    return reference_root() is not None


def load_reference():
    """Returns the reference ``retinanet`` package (anchors, box_utils, losses, models, layers)."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not found (RN_REFERENCE_ROOT, /root/reference, baseline/_ref)")
    import torch

    if "torchvision.models.utils" not in sys.modules:
        stub = types.ModuleType("torchvision.models.utils")
        stub.load_state_dict_from_url = torch.hub.load_state_dict_from_url
        sys.modules["torchvision.models.utils"] = stub
    if root not in sys.path:
        sys.path.insert(0, root)
    import retinanet  # noqa: F401
    import retinanet.anchors, retinanet.box_utils, retinanet.layers, retinanet.losses, retinanet.models  # noqa: F401,E401

    return retinanet


def reference_step(ref, cls_preds, bbox_preds, targets, padded_hw, im_szs, num_classes, backward=True,
                   score_thres=0.05, nms_thres=0.5, detections_per_img=100, strides=(8, 16, 32, 64, 128)):
    """One pass of the path through the reference's own entry points, on whatever device the tensors live:
    ``AnchorGenerator.forward`` (anchors.py:199) -> ``RetinaNetLosses.forward`` (losses.py:113) + backward ->
    ``Retinanet.process_detections`` (models.py:160).  Returns (loss dict, detections)."""
    import torch

    n = cls_preds.shape[0]
    dev = cls_preds.device
    gen = ref.anchors.AnchorGenerator().to(dev)
    h, w = padded_hw
    fmaps = [torch.empty((n, 1, -(-h // s), -(-w // s)), device=dev) for s in strides]
    anchors = gen(SimpleNamespace(image_sizes=list(im_szs)), fmaps)
    x = cls_preds.detach().clone().requires_grad_(backward)
    b = bbox_preds.detach().clone().requires_grad_(backward)
    losses = ref.losses.RetinaNetLosses(num_classes)
    out = losses(targets, {"cls_preds": x, "bbox_preds": b}, anchors)
    if backward:
        (out["classification_loss"] + out["regression_loss"]).backward()
    stub = SimpleNamespace(score_thres=score_thres, nms_thres=nms_thres, detections_per_img=detections_per_img)
    with torch.no_grad():
        dets = ref.models.Retinanet.process_detections(stub, {"cls_preds": x.detach(), "bbox_preds": b.detach().clone()},
                                                       anchors, list(im_szs))
    return out, dets
