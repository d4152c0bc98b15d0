"""Import the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

Thin alias of ``baseline/reference.py`` (kept for the fixture generator ``tests/golden/make_golden.py`` and the CPU
tests): the reference is looked up at ``$RN_REFERENCE_ROOT``, ``/root/reference`` (build container) and
``baseline/_ref`` (the offline install that travels to the GPU box), and imported behind the one-module
``torchvision.models.utils`` stub.  No reference file is touched or copied into the repository's history.
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from baseline.reference import load_reference, reference_available, reference_root  # noqa: E402,F401

REFERENCE_ROOT = reference_root()
