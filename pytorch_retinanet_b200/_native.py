"""ctypes binding of the C ABI declared in ``include/retinanet_b200.h``.

The shared library is built in-tree (``pytorch_retinanet_b200/lib/librn_b200.so``) by
``pytorch_retinanet_b200/build.py`` / ``__graft_entry__.build()``.  There is NO fallback: if the
library is missing and cannot be built, or a tensor is not a contiguous CUDA tensor of the expected
dtype, the call raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

from . import build as _build

_c = ctypes
_vp, _i32, _i64, _f32, _f64, _sz = _c.c_void_p, _c.c_int32, _c.c_int64, _c.c_float, _c.c_double, _c.c_size_t

RN_MAX_PEERS = 16


class RnExchange(_c.Structure):
    """``rn_exchange_t`` of include/retinanet_b200.h."""
    _fields_ = [("peers", _vp * RN_MAX_PEERS), ("rank", _i32), ("world", _i32)]


_xp = _c.POINTER(RnExchange)

# name -> (restype, argtypes); must list every symbol declared in include/retinanet_b200.h
SIGNATURES = {
    "rn_comm_bytes": (_sz, []),
    "rn_comm_alloc": (_c.c_int, [_c.POINTER(_vp)]),
    "rn_comm_free": (_c.c_int, [_vp]),
    "rn_comm_export": (_c.c_int, [_vp, _vp]),
    "rn_comm_import": (_c.c_int, [_vp, _c.POINTER(_vp)]),
    "rn_comm_unmap": (_c.c_int, [_vp]),
    "rn_comm_error": (_c.c_int, [_vp, _c.POINTER(_i32)]),
    "rn_exchange_total": (_c.c_int, [_vp, _xp, _vp]),
    "rn_abi_version": (_c.c_int, []),
    "rn_last_error": (_c.c_char_p, []),
    "rn_launch_count": (_c.c_uint64, []),
    "rn_anchor_grid": (_c.c_int, [_vp, _vp, _c.c_int, _f64, _vp, _i64, _vp]),
    "rn_pack_targets": (_c.c_int, [_vp, _vp, _vp, _c.c_int, _vp, _vp, _vp, _vp, _vp]),
    "rn_match": (_c.c_int, [_vp, _i64, _i64, _vp, _vp, _vp, _c.c_int, _i64, _f32, _f32, _vp, _vp, _vp, _vp]),
    "rn_encode": (_c.c_int, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "rn_decode": (_c.c_int, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "rn_loss_workspace_bytes": (_sz, [_c.c_int, _i64, _c.c_int]),
    "rn_loss": (_c.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _c.c_int, _i64, _c.c_int, _f32, _f32, _f32, _vp,
                           _f32, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _xp]),
    "rn_train_loss_workspace_bytes": (_sz, [_c.c_int, _i64, _c.c_int]),
    "rn_train_loss": (_c.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _c.c_int, _i64, _i64, _c.c_int, _f32, _f32, _f32, _f32,
                                 _f32, _vp, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _xp]),
    "rn_train_detect_workspace_bytes": (_sz, [_c.c_int, _i64, _c.c_int, _i64, _c.c_int]),
    "rn_train_detect": (_c.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _c.c_int, _i64, _i64, _c.c_int, _f32, _f32, _f32, _f32, _f32,
                                   _vp, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _f64, _c.c_int, _c.c_int, _vp, _c.c_int,
                                   _i64, _vp, _vp, _vp, _vp, _vp, _vp, _c.c_int, _vp, _sz, _vp, _xp, _c.c_int]),
    "rn_dense_loss_workspace_bytes": (_sz, []),
    "rn_focal_loss_dense": (_c.c_int, [_vp, _vp, _i64, _f32, _f32, _vp, _vp, _vp, _sz, _vp]),
    "rn_smooth_l1_dense": (_c.c_int, [_vp, _vp, _i64, _f32, _vp, _vp, _vp, _sz, _vp]),
    "rn_scale_by_device_scalar": (_c.c_int, [_vp, _i64, _vp, _vp]),
    "rn_postprocess_workspace_bytes": (_sz, [_c.c_int, _i64, _c.c_int, _i64, _c.c_int]),
    "rn_postprocess": (_c.c_int, [_vp, _vp, _vp, _i64, _vp, _c.c_int, _i64, _c.c_int, _f32, _f64, _c.c_int, _vp,
                                  _c.c_int, _vp, _c.c_int, _c.c_int, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp, _c.c_int]),
    "rn_loss_levels_workspace_bytes": (_sz, [_c.c_int, _vp, _c.c_int]),
    "rn_loss_levels": (_c.c_int, [_vp, _vp, _vp, _c.c_int, _vp, _i64, _vp, _vp, _vp, _vp, _c.c_int, _i64, _c.c_int, _f32,
                                  _f32, _f32, _vp, _f32, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _xp]),
    "rn_postprocess_levels_workspace_bytes": (_sz, [_c.c_int, _i64, _c.c_int, _i64, _c.c_int]),
    "rn_postprocess_levels": (_c.c_int, [_vp, _vp, _vp, _c.c_int, _vp, _i64, _vp, _c.c_int, _i64, _c.c_int, _f32, _f64,
                                         _c.c_int, _vp, _c.c_int, _c.c_int, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp, _c.c_int]),
    "rn_nms_segments": (_c.c_int, [_vp, _vp, _c.c_int, _i64, _f64, _vp, _vp, _sz, _vp]),
    "rn_loss_set_math_mode": (_c.c_int, [_c.c_int]),
}

_lib: Optional[ctypes.CDLL] = None


class NativeError(RuntimeError):
    pass


def lib_path() -> str:
    return _build.LIB_PATH


def header_abi_version() -> int:
    """``RN_ABI_VERSION`` as declared in include/retinanet_b200.h — the version these bindings were written for."""
    import re
    with open(os.path.join(_build.ROOT, "include", "retinanet_b200.h")) as f:
        m = re.search(r"#define\s+RN_ABI_VERSION\s+(\d+)", f.read())
    if not m:
        raise NativeError("retinanet_b200: RN_ABI_VERSION not found in include/retinanet_b200.h")
    return int(m.group(1))


def load() -> ctypes.CDLL:
    """Loads (building first if the sources are newer and nvcc is present) the CUDA library."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if _build.needs_build():
        try:
            _build.build()
        except Exception as e:  # no nvcc on this box and no prebuilt library -> fail loudly
            if not os.path.exists(path):
                raise NativeError(
                    f"retinanet_b200: CUDA library {path} is missing and could not be built ({e}). "
                    "Run `python __graft_entry__.py` (build()) on a machine with nvcc; there is no CPU fallback."
                ) from e
            import warnings                     # an older library exists: say so — the checks below catch ABI drift
            warnings.warn(f"retinanet_b200: sources are newer than {path} but the rebuild failed ({e}); using the "
                          "existing library", RuntimeWarning)
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:              # stale library: fail loudly
            raise NativeError(f"retinanet_b200: {path} does not export {name} — stale build, run build()") from e
        fn.restype, fn.argtypes = res, args
    want = header_abi_version()
    if lib.rn_abi_version() != want:
        raise NativeError(f"retinanet_b200: ABI version mismatch (library {lib.rn_abi_version()}, header {want}): "
                          "stale build, run build()")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().rn_last_error().decode("utf-8", "replace")
        raise NativeError(f"{what} failed (code {rc}): {msg}")


def ptr(t: Optional[torch.Tensor], dtype: Optional[torch.dtype] = None, what: str = "tensor") -> Optional[int]:
    """Device pointer of a contiguous CUDA tensor (None -> NULL).  Raises on CPU tensors: no fallback."""
    if t is None:
        return None
    if not t.is_cuda:
        raise NativeError(f"retinanet_b200: {what} must be a CUDA tensor (got device {t.device}); there is no CPU path")
    if dtype is not None and t.dtype != dtype:
        raise NativeError(f"retinanet_b200: {what} must have dtype {dtype} (got {t.dtype})")
    if not t.is_contiguous():
        raise NativeError(f"retinanet_b200: {what} must be contiguous")
    return t.data_ptr()


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class on_device:
    """``with on_device(dev):`` — like ``torch.cuda.device(dev)`` but free when ``dev`` already is the current
    device (the common case: one process per GPU); the host path of a step is latency critical."""

    __slots__ = ("ctx",)

    def __init__(self, dev: torch.device):
        idx = dev.index
        self.ctx = None if (idx is None or idx == torch.cuda.current_device()) else torch.cuda.device(dev)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)


def host_floats(vals) -> ctypes.Array:
    return (_f32 * len(vals))(*[float(v) for v in vals])
