"""WHDR (weighted human disagreement rate) of reflectance images on the device.

Mirror of /root/reference/training/layers/whdr_layer.py: ``whdr`` (:253-287), ``get_comparisons_from_blob``
(:216-237), ``_extract_valid_comparisons_with_actual_size`` (:240-251) and the batch mean of
``WhdrLayer.forward`` (:71-88).  The comparison blob is the one createNumpyArrayWithComparisonsForIIW.py:616-649
writes: ``[max+1, 6]`` float64 rows ``(x1, y1, x2, y2, darker, weight)`` with relative coordinates, NaN padding
and a last row ``(count, file name, 0)``.  All arithmetic runs in ``rf_whdr_f32``; there is no CPU fallback.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import _native, device as dev


def _as_blob(comparisons: torch.Tensor) -> torch.Tensor:
    if comparisons.dim() == 4 and comparisons.shape[2] == 1:  # the Caffe blob layout [n, max+1, 1, 6]
        comparisons = comparisons[:, :, 0, :]
    if comparisons.dim() != 3 or comparisons.shape[2] != 6 or comparisons.shape[1] < 1:
        raise ValueError("comparisons must have shape [n, max_comparisons + 1, 6]")
    if comparisons.dtype != torch.float64 or not comparisons.is_cuda:
        raise ValueError("comparisons must be a float64 CUDA tensor")
    return comparisons.contiguous()


def whdr_sums_device(reflectance: torch.Tensor, comparisons: torch.Tensor, delta: float = 0.1,
                     pixel_coords: bool = False) -> torch.Tensor:
    """``[n, 2]`` float64 ``(error_sum, weight_sum)`` per image.

    reflectance: float32 CUDA ``[n, h, w]`` (the CNN's intensity) or ``[n, c, h, w]`` with c in {1, 3};
    comparisons: float64 CUDA ``[n, max+1, 6]`` (or ``[n, max+1, 1, 6]``).  Raises ``IndexError`` when a
    comparison lies outside the image (as numpy indexing does) and ``ValueError`` for a broken count row.
    """
    if reflectance.dtype != torch.float32 or not reflectance.is_cuda:
        raise ValueError("reflectance must be a float32 CUDA tensor")
    if reflectance.dim() == 3:
        reflectance = reflectance[:, None]
    if reflectance.dim() != 4:
        raise ValueError("reflectance must have shape [n, h, w] or [n, c, h, w]")
    n, c, h, w = reflectance.shape
    if c not in (1, 3):
        raise Exception("Expecting 1 or 3 channels to compute lightness!")
    blob = _as_blob(comparisons)
    if blob.shape[0] != n:
        raise ValueError("one comparison blob per image expected")
    if blob.device != reflectance.device:
        raise ValueError("reflectance and comparisons must be on the same device")
    reflectance = reflectance.contiguous()
    out = torch.empty((n, 2), dtype=torch.float64, device=reflectance.device)
    bad = torch.zeros(1, dtype=torch.int32, device=reflectance.device)
    with torch.cuda.device(reflectance.device):
        dev.bind_device(reflectance.device)
        _native.check(_native.lib().rf_whdr_f32(
            dev.ptr(reflectance), c, n, h, w, dev.ptr(blob), blob.shape[1] - 1, float(delta),
            _native.RF_WHDR_PIXEL_COORDS if pixel_coords else 0, dev.ptr(out), dev.ptr(bad), dev.stream_ptr()))
    flag = int(bad.item())
    if flag & 1:
        raise ValueError("comparison blob: the count in the last row is not in [0, max_comparisons]")
    if flag & 2:
        raise IndexError("comparison coordinates are out of bounds for the reflectance image")
    return out


def whdr_device(reflectance: torch.Tensor, comparisons: torch.Tensor, delta: float = 0.1,
                pixel_coords: bool = False) -> torch.Tensor:
    """Per-image WHDR ``[n]`` float64: error_sum / weight_sum, 0 where there are no comparisons."""
    s = whdr_sums_device(reflectance, comparisons, delta, pixel_coords)
    return torch.where(s[:, 1] != 0, s[:, 0] / torch.where(s[:, 1] != 0, s[:, 1], torch.ones_like(s[:, 1])),
                       torch.zeros_like(s[:, 0]))


def reduce_mean(per_image: torch.Tensor, group=None) -> Tuple[float, int]:
    """``(mean, count)`` of per-image WHDRs over every rank of the process group (one all-reduce of
    ``[sum, count]`` in float64; without torch.distributed the local mean).  An empty shard contributes (0, 0)."""
    acc = torch.zeros(2, dtype=torch.float64, device=per_image.device)
    acc[0] = per_image.to(torch.float64).sum()
    acc[1] = float(per_image.numel())
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.all_reduce(acc, group=group)
    total, count = acc.tolist()
    return (total / count if count else 0.0), int(count)


def mean_whdr(reflectance: torch.Tensor, comparisons: torch.Tensor, delta: float = 0.1, group=None) -> Tuple[float, int]:
    """Mean of the per-image WHDRs (what WhdrLayer.forward reports for a batch, whdr_layer.py:71-88), taken over
    every rank's images when ``torch.distributed`` is initialised."""
    return reduce_mean(whdr_device(reflectance, comparisons, delta), group)


# ---- numpy-facing mirror of the reference functions -------------------------------------------------------------
def get_comparisons_from_blob(comp_blob: np.ndarray, height: int, width: int, delta: float,
                              ground_truth_albedo=None):
    """``([num_comparisons, 6] array in pixel coordinates, file name)`` -- whdr_layer.py:216-237 (host-side view
    of the blob; the device kernel applies the same scaling itself)."""
    file_name = comp_blob[-1, 0, 1]
    comparisons = comp_blob[:, 0, :]
    num = int(comparisons[-1, 0])
    res = comparisons[:num, :].copy()
    res[:, [0, 2]] = (res[:, [0, 2]] * width).astype(int)
    res[:, [1, 3]] = (res[:, [1, 3]] * height).astype(int)
    return res, file_name


def whdr(reflectance: np.ndarray, comparisons: np.ndarray, delta: float, device=None) -> float:
    """``whdr(reflectance [c, h, w], comparisons [num, 6] in pixel coordinates, delta)`` -- whdr_layer.py:253."""
    reflectance = np.asarray(reflectance)
    if reflectance.ndim != 3:
        raise ValueError("Expects a reflectance image of shape [c, h, w]")
    comparisons = np.asarray(comparisons, dtype=np.float64).reshape(-1, 6)
    num = comparisons.shape[0]
    blob = np.full((1, num + 1, 6), np.nan)
    blob[0, :num] = comparisons
    blob[0, num, :3] = (num, 0.0, 0.0)
    d = dev.bind_device(device)
    r = torch.from_numpy(np.ascontiguousarray(reflectance, dtype=np.float32))[None].to(d)
    return float(whdr_device(r, torch.from_numpy(blob).to(d), delta, pixel_coords=True)[0].item())
