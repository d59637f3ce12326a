"""Device-resident CNN -> filter pipelines and batch sharding.

The reference chains its stages through PNG files: ``decompose_with_trained_CNN.py`` writes
``<base>-r.png`` (``trunc(r * 255)``, one channel), ``filter_reflectance.py`` reads it back with
``cv2.imread`` (three equal channels) and filters it, and the "3x GF" configuration feeds each
output PNG back in.  The pipelines here keep exactly those quantisation points --
truncation after the CNN, round-half-even uint8 after every filter application -- but keep the
data on the device; a gray image is carried as one plane and expanded to three equal channels
only when it leaves.

Images are independent, so a batch is split into contiguous shards, one per GPU / process, with
no communication on the pixel path.  The only collective is the optional
all-reduce of :func:`aggregate_stats`.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import _native, cnn, device as dev, filters


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard ``[lo, hi)`` of ``n_items`` for ``rank`` of ``world``: item ``i`` goes to
    rank ``i * world // n_items``."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    lo = (rank * n_items + world - 1) // world
    hi = ((rank + 1) * n_items + world - 1) // world
    return lo, hi


class Pipeline(object):
    """CNN -> joint bilateral / guided filter on one device."""

    def __init__(self, net: Optional[cnn.Net] = None, device=None):
        self.net = net if net is not None else cnn.default_net(device)
        self.device = self.net.device
        self._host_ctx = {}

    # ---- device-resident stages ------------------------------------------------------------
    def reflectance_u8(self, images: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``uint8[N,H,W,3]`` BGR -> ``uint8[N,H,W]``: the bytes of ``<base>-r.png``."""
        return self.net.forward_device(images, want_f32=False, want_u8=True, out_u8=out)[1]

    def cnn_bf(self, images: torch.Tensor, sigma_color: float = 20.0, sigma_spatial: float = 22.0,
               out: Optional[torch.Tensor] = None, scratch: Optional[torch.Tensor] = None) -> torch.Tensor:
        """BF(CNN, CNN): the CNN reflectance filtered with itself as guidance.  Returns the gray
        plane ``uint8[N,H,W]`` (the reference's PNG has this value in all three channels).
        ``scratch``: optional ``uint8[N,H,W]`` buffer for the intermediate reflectance."""
        r = self.reflectance_u8(images, out=scratch)
        return filters.joint_bilateral_device(r, r, sigma_color, sigma_spatial, d=-1,
                                              gray_replicated=True, out=out)

    def cnn_gf(self, images: torch.Tensor, guides: torch.Tensor, sigma_color: float = 3.0,
               sigma_spatial: float = 45.0, iterations: int = 3, out: Optional[torch.Tensor] = None,
               scratch: Optional[torch.Tensor] = None) -> torch.Tensor:
        """GF(CNN, guide) applied ``iterations`` times with uint8 re-quantisation between
        applications (three CLI invocations in the reference).  Returns ``uint8[N,H,W]``.
        ``out`` / ``scratch``: optional ``uint8[N,H,W]`` buffers for the result / the intermediate reflectance."""
        if iterations < 1:
            raise ValueError("iterations must be >= 1")
        cur = self.reflectance_u8(images, out=scratch)
        return filters.guided_device(guides, cur, int(sigma_spatial), sigma_color, out=out, iterations=iterations)

    # ---- host buffers in, host buffers out ----------------------------------------------------
    def run_host(self, kind: str, images: torch.Tensor, out: torch.Tensor,
                 guides: Optional[torch.Tensor] = None, chunk: int = 32, n_streams: int = 4,
                 **params) -> None:
        """End-to-end over pinned HOST tensors: ``images`` ``uint8[N,H,W,3]`` (and ``guides``) are
        copied to the device chunk by chunk, run through ``cnn_bf`` / ``cnn_gf``, and the gray
        results copied back into ``out`` ``uint8[N,H,W]``; copies and kernels of different chunks
        overlap on ``n_streams`` streams.  Returns after everything has landed in ``out``."""
        if kind not in ("cnn_bf", "cnn_gf"):
            raise ValueError("kind must be 'cnn_bf' or 'cnn_gf'")
        if kind == "cnn_gf" and guides is None:
            raise ValueError("cnn_gf needs guides")
        n, h, w = images.shape[0], images.shape[1], images.shape[2]
        dev.bind_device(self.device)
        # streams and per-stream device buffers are created once per (shape, chunk) and reused: fresh
        # streams / allocations on every call cost tens of milliseconds of allocator traffic
        key = (kind, chunk, h, w, max(1, n_streams))
        ctx = self._host_ctx.get(key)
        if ctx is None:
            if len(self._host_ctx) > 8:
                self._host_ctx.clear()
            ctx = {"streams": [torch.cuda.Stream(device=self.device) for _ in range(max(1, n_streams))], "bufs": []}
            for _ in ctx["streams"]:
                b = {"img": torch.empty((chunk, h, w, 3), dtype=torch.uint8, device=self.device),
                     "r8": torch.empty((chunk, h, w), dtype=torch.uint8, device=self.device),
                     "out": torch.empty((chunk, h, w), dtype=torch.uint8, device=self.device)}
                if kind == "cnn_gf":
                    b["guide"] = torch.empty((chunk, h, w, 3), dtype=torch.uint8, device=self.device)
                ctx["bufs"].append(b)
            self._host_ctx[key] = ctx
        streams = ctx["streams"]
        cur = torch.cuda.current_stream()
        for s in streams:
            s.wait_stream(cur)
        # the first chunks are small so that the kernels start after a short first copy (nothing overlaps the first
        # host-to-device transfer); then full chunks
        bounds, lo = [], 0
        for size in (max(1, chunk // 8), max(1, chunk // 4), max(1, chunk // 2)):
            if lo + size < n and size < chunk:
                bounds.append((lo, lo + size))
                lo += size
        while lo < n:
            bounds.append((lo, min(n, lo + chunk)))
            lo += chunk
        for ci, (lo, hi) in enumerate(bounds):
            m = hi - lo
            s = streams[ci % len(streams)]
            b = ctx["bufs"][ci % len(streams)]
            with torch.cuda.stream(s):
                d_img = b["img"][:m]
                d_img.copy_(images[lo:hi], non_blocking=True)
                if kind == "cnn_bf":
                    res = self.cnn_bf(d_img, out=b["out"][:m], scratch=b["r8"][:m], **params)
                else:
                    d_gd = b["guide"][:m]
                    d_gd.copy_(guides[lo:hi], non_blocking=True)
                    res = self.cnn_gf(d_img, d_gd, out=b["out"][:m], scratch=b["r8"][:m], **params)
                out[lo:hi].copy_(res, non_blocking=True)
        for s in streams:
            cur.wait_stream(s)
        cur.synchronize()


def aggregate_stats(inp: torch.Tensor, out: torch.Tensor, group=None) -> np.ndarray:
    """``[n_bytes, sum(out), sum(out^2), sum|out-in|]`` over this rank's tensors, summed over the
    process group if ``torch.distributed`` is initialised -- the one collective of a multi-GPU
    run (an all-reduce of four doubles)."""
    dev.check_u8_cuda(inp, "inp")
    dev.check_u8_cuda(out, "out")
    if inp.numel() != out.numel():
        raise ValueError("inp and out must have the same number of bytes")
    stats = torch.zeros(4, dtype=torch.float64, device=out.device)
    with torch.cuda.device(out.device):
        dev.bind_device(out.device)
        _native.check(_native.lib().rf_accumulate_stats_u8(dev.ptr(inp), dev.ptr(out), out.numel(),
                                                           dev.ptr(stats), dev.stream_ptr()))
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.all_reduce(stats, group=group)
    return stats.cpu().numpy()
