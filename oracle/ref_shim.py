"""Import the UNMODIFIED reference (``/root/reference``) in the build container.

TEST INFRASTRUCTURE ONLY.  ``/root/reference`` does not exist on the GPU box, so
this module is used solely by ``tests/golden/make_golden.py`` (fixture
generation) and by the CPU tests that are skipped when the tree is absent.

The single import obstacle is ``retinanet/backbone.py:6``
(``from torchvision.models.utils import load_state_dict_from_url`` — a module
removed from current torchvision).  A stub module is registered in
``sys.modules`` before the import; no reference file is touched or copied.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("RN_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "retinanet"))


def load_reference():
    """Returns the reference ``retinanet`` package (anchors, box_utils, losses, models)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import torch

    if "torchvision.models.utils" not in sys.modules:
        stub = types.ModuleType("torchvision.models.utils")
        stub.load_state_dict_from_url = torch.hub.load_state_dict_from_url
        sys.modules["torchvision.models.utils"] = stub
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import retinanet  # noqa: F401
    import retinanet.anchors, retinanet.box_utils, retinanet.losses, retinanet.models  # noqa: F401,E401

    return retinanet
