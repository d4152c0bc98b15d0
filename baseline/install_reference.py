#!/usr/bin/env python
"""Installs the UNMODIFIED reference (``/root/reference``) into ``baseline/_ref`` (git-ignored, travels to the GPU box).

The reference ships neither ``setup.py`` nor ``pyproject.toml``, so the contract's command

    python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference

stops with "Neither 'setup.py' nor 'pyproject.toml' found".  This script does what the contract allows for that case:
it copies the tree to a scratch directory under /tmp (``/root/reference`` is read-only), adds the ONE missing file —
a five-line ``setup.py`` that names the ``retinanet`` package, nothing else — and runs the same pip command on the
copy with ``--no-deps`` (the reference's pinned requirements are its trainer's: pytorch-lightning, albumentations,
pycocotools … none is used by the path, none is in the wheelhouse).  The installed files are byte-identical to
``/root/reference/retinanet/*.py`` (checked below); nothing of the reference enters the git history.

Used by: ``bench.py --impl reference`` / ``cpu_baseline`` (kind "reference"), the real-model tests, ``oracle/ref_shim``.
"""
import filecmp
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TARGET = os.path.join(ROOT, "baseline", "_ref")
SOURCE = os.environ.get("RN_REFERENCE_SOURCE", "/root/reference")

SETUP_PY = '''from setuptools import setup
setup(name="pytorch_retinanet_reference", version="0.0.0", packages=["retinanet"],
      description="benihime91/pytorch_retinanet, unmodified (retinanet package only)")
'''


def install(force: bool = False) -> str:
    if not os.path.isdir(os.path.join(SOURCE, "retinanet")):
        raise RuntimeError(f"no reference tree at {SOURCE}")
    if os.path.isdir(os.path.join(TARGET, "retinanet")) and not force and verify(quiet=True):
        return TARGET
    tmp = tempfile.mkdtemp(prefix="rn_ref_src_")
    try:
        src = os.path.join(tmp, "reference")
        shutil.copytree(SOURCE, src, ignore=shutil.ignore_patterns("*.ipynb", ".git"))
        with open(os.path.join(src, "setup.py"), "w") as f:
            f.write(SETUP_PY)
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
               "/opt/wheelhouse", "--target", TARGET, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("pip install of the reference failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    if not verify():
        raise RuntimeError("installed reference differs from the source tree")
    return TARGET


def verify(quiet: bool = False) -> bool:
    """Every installed ``retinanet/*.py`` is byte-identical to the source tree's."""
    src, dst = os.path.join(SOURCE, "retinanet"), os.path.join(TARGET, "retinanet")
    names = sorted(n for n in os.listdir(src) if n.endswith(".py"))
    _, mismatch, errors = filecmp.cmpfiles(src, dst, names, shallow=False)
    if (mismatch or errors) and not quiet:
        print("reference install mismatch:", mismatch, errors, file=sys.stderr)
    return not (mismatch or errors)


if __name__ == "__main__":
    print("installed", install(force="--force" in sys.argv))
