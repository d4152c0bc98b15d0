"""Builds the in-tree CUDA library ``pytorch_retinanet_b200/lib/librn_b200.so`` with plain nvcc for
sm_100a (no torch C++ extension: the C ABI in include/retinanet_b200.h has no torch types)."""
from __future__ import annotations

import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "librn_b200%s.so" % os.environ.get("RN_LIB_SUFFIX", "_dbg" if os.environ.get("RN_LAZY_TIMING") else ""))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = (["-DRN_LAZY_TIMING=1"] if os.environ.get("RN_LAZY_TIMING") else []) + os.environ.get("RN_EXTRA_DEFS", "").split()
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", LIB_PATH] + sources()
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(LIB_DIR, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed (see {log})")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
