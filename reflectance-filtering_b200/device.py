"""Device plumbing: torch owns memory, streams and copies; the kernels are librf_b200's.

Everything here fails loudly when CUDA or the built library is missing (no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Dict, Tuple

import numpy as np
import torch

from . import _native


def require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("reflectance-filtering_b200 needs a CUDA device (sm_100a); "
                           "torch.cuda.is_available() is False and there is no CPU fallback")
    _native.lib()


def bind_device(device: torch.device | int | None = None) -> torch.device:
    """Make ``device`` current for torch and for the library's own CUDA runtime instance."""
    require_cuda()
    if device is None:
        idx = torch.cuda.current_device()
    else:
        idx = torch.device(device).index if not isinstance(device, int) else device
        if idx is None:
            idx = torch.cuda.current_device()
    torch.cuda.set_device(idx)
    _native.check(_native.lib().rf_set_device(idx))
    return torch.device("cuda", idx)


def ptr(t: torch.Tensor) -> C.c_void_p:
    return C.c_void_p(t.data_ptr())


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def check_u8_cuda(t: torch.Tensor, name: str) -> None:
    if not isinstance(t, torch.Tensor) or t.dtype != torch.uint8 or not t.is_cuda:
        raise TypeError("%s must be a CUDA uint8 tensor" % name)
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)


class PinnedStaging:
    """Per-thread cache of pinned host buffers keyed by (tag, nbytes): host arrays go through
    these so the H2D / D2H copies are true async DMA transfers."""

    def __init__(self):
        self._local = threading.local()

    def get(self, tag: str, shape: Tuple[int, ...], dtype=torch.uint8) -> torch.Tensor:
        cache: Dict = self._local.__dict__.setdefault("cache", {})
        key = (tag, tuple(shape), dtype)
        buf = cache.get(key)
        if buf is None:
            if len(cache) > 32:
                cache.clear()
            buf = torch.empty(shape, dtype=dtype, pin_memory=True)
            cache[key] = buf
        return buf


staging = PinnedStaging()


def to_device(arr: np.ndarray, tag: str) -> torch.Tensor:
    """numpy (host, pageable) -> pinned staging -> device tensor on the current stream."""
    arr = np.ascontiguousarray(arr)
    pin = staging.get(tag, arr.shape, torch.from_numpy(arr).dtype)
    pin.numpy()[...] = arr
    return pin.to("cuda", non_blocking=True)


def to_host(t: torch.Tensor, tag: str) -> np.ndarray:
    """device tensor -> pinned staging -> fresh numpy array (synchronises the current stream)."""
    pin = staging.get(tag, tuple(t.shape), t.dtype)
    pin.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return pin.numpy().copy()
