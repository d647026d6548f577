"""Device hand-off: PyTorch owns device memory and streams, nothing else.

Every kernel is reached through the C-ABI with raw `data_ptr()`s and the raw
stream handle; torch tensors never cross the library boundary.
"""
from __future__ import annotations

import numpy as np

from . import _lib

_torch = None


def torch():
    global _torch
    if _torch is None:
        import torch as _t
        _torch = _t
    return _torch


def require_cuda():
    t = torch()
    if not t.cuda.is_available():
        raise _lib.OisatError(
            "no CUDA device: oisatgmi_b200 runs its hot path on the GPU only (no CPU fallback)")
    _lib.lib()
    return t


def device():
    t = require_cuda()
    return t.device("cuda", t.cuda.current_device())


def stream():
    return torch().cuda.current_stream().cuda_stream


_NP2CODE = {np.dtype(np.float16): _lib.F16, np.dtype(np.float32): _lib.F32,
            np.dtype(np.float64): _lib.F64, np.dtype(np.uint8): _lib.U8,
            np.dtype(np.int32): _lib.I32}


def dtype_code(arr) -> int:
    dt = np.dtype(arr.dtype) if isinstance(arr, np.ndarray) else _torch_np_dtype(arr)
    return _NP2CODE[dt]


def _torch_np_dtype(t):
    tt = torch()
    return {tt.float16: np.dtype(np.float16), tt.float32: np.dtype(np.float32),
            tt.float64: np.dtype(np.float64), tt.uint8: np.dtype(np.uint8),
            tt.int32: np.dtype(np.int32)}[t.dtype]


def native_float(arr: np.ndarray) -> np.ndarray:
    """Reader arrays keep their dtype when it is f16/f32/f64; anything else
    (ints, bools, object) is widened to float64 exactly as numpy would on the
    first multiply with the float64 mask (interpolator.py:127,163)."""
    arr = np.asarray(arr)
    if arr.dtype in (np.float16, np.float32, np.float64):
        return arr
    return arr.astype(np.float64)


_keep = None   # a list while a recording wants every allocation kept alive (see keep_allocations)


class keep_allocations:
    """While active, every tensor the helpers below create is also appended to `store`: the
    launches recorded by `_lib.recording` refer to raw addresses, which must stay valid for as
    long as the recording is replayed."""

    def __init__(self, store):
        self.store = store

    def __enter__(self):
        global _keep
        self._saved, _keep = _keep, self.store
        return self.store

    def __exit__(self, *exc):
        global _keep
        _keep = self._saved
        return False


def _kept(t):
    if _keep is not None:
        _keep.append(t)
    return t


def to_device(arr, dtype=None, pin=False):
    """numpy -> device tensor (contiguous, dtype preserved unless given)."""
    t = require_cuda()
    a = np.ascontiguousarray(arr if dtype is None else np.asarray(arr, dtype=dtype))
    h = t.from_numpy(a)
    if pin:
        h = h.pin_memory()
    return _kept(h.to(device(), non_blocking=pin))


def empty(shape, dtype="float64"):
    t = require_cuda()
    return _kept(t.empty(shape, dtype=getattr(t, dtype), device=device()))


def full(shape, value, dtype="float64"):
    t = require_cuda()
    return _kept(t.full(shape, value, dtype=getattr(t, dtype), device=device()))


def zeros(shape, dtype="float64"):
    t = require_cuda()
    return _kept(t.zeros(shape, dtype=getattr(t, dtype), device=device()))


def ptr(t):
    return None if t is None else t.data_ptr()


def to_host(t) -> np.ndarray:
    return t.detach().cpu().numpy()
