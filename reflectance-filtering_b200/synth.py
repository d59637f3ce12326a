"""Deterministic synthetic inputs.  All images are ``uint8[H,W,3]`` BGR.

``seed = 1000 * config_id + image_index`` with ``np.random.default_rng(seed)``.

* ``natural``: piecewise-smooth colour field (3 low-frequency sinusoids per channel,
  amplitude 100 around 127) + Voronoi-cell albedo offsets (64 cells, +-40) + Gaussian
  noise sigma 12, clipped.
* ``flat``: piecewise-constant Voronoi palette (64 cells, colours U[30,230]) + noise
  sigma 2 -- stands in for the 'flat' guidance image of the 3xGF configuration.
* ``stress``: i.i.d. U{0..255}.
"""
from __future__ import annotations

import numpy as np


def _voronoi_labels(rng, h, w, cells):
    cy = rng.uniform(0, h, cells).astype(np.float32)
    cx = rng.uniform(0, w, cells).astype(np.float32)
    # nearest seed on a coarse grid, then upsample: exact Voronoi is not the point, speed is
    step = 4 if max(h, w) > 256 else 1
    ys = np.arange(0, h, step, dtype=np.float32)[:, None, None]
    xs = np.arange(0, w, step, dtype=np.float32)[None, :, None]
    d = (ys - cy[None, None, :]) ** 2 + (xs - cx[None, None, :]) ** 2
    lab = np.argmin(d, axis=2)
    if step > 1:
        lab = np.repeat(np.repeat(lab, step, axis=0), step, axis=1)[:h, :w]
    return lab


def natural(h: int, w: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    yy = np.arange(h, dtype=np.float32)[:, None] / max(h, 1)
    xx = np.arange(w, dtype=np.float32)[None, :] / max(w, 1)
    img = np.empty((h, w, 3), np.float32)
    for c in range(3):
        f = np.zeros((h, w), np.float32)
        for _ in range(3):
            fy, fx = rng.uniform(0.5, 3.0, 2)
            ph = rng.uniform(0, 2 * np.pi)
            f += np.sin(2 * np.pi * (fy * yy + fx * xx) + ph).astype(np.float32)
        img[:, :, c] = 127.0 + (100.0 / 3.0) * f
    lab = _voronoi_labels(rng, h, w, 64)
    offs = rng.uniform(-40, 40, (64, 3)).astype(np.float32)
    img += offs[lab]
    img += rng.normal(0, 12, img.shape).astype(np.float32)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def flat(h: int, w: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    lab = _voronoi_labels(rng, h, w, 64)
    pal = rng.uniform(30, 230, (64, 3)).astype(np.float32)
    img = pal[lab] + rng.normal(0, 2, (h, w, 3)).astype(np.float32)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def stress(h: int, w: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)


GENERATORS = {"natural": natural, "flat": flat, "stress": stress}


def batch(kind: str, n: int, h: int, w: int, config_id: int, start: int = 0) -> np.ndarray:
    """``uint8[n,h,w,3]`` with per-image seeds ``1000*config_id + start + i``."""
    gen = GENERATORS[kind]
    return np.stack([gen(h, w, 1000 * config_id + start + i) for i in range(n)])


def comparisons(n: int, max_comparisons: int, seed: int, min_count: int = 0) -> np.ndarray:
    """Synthetic IIW-style comparison blobs ``[n, max_comparisons + 1, 1, 6]`` float64 in the layout of the
    reference's createNumpyArrayWithComparisonsForIIW.py:616-649: rows ``(x1, y1, x2, y2, darker, weight)`` with
    relative coordinates in [0, 1), darker in {0, 1, 2}, weights in (0, 3]; NaN padding; the last row holds
    ``(count, file name, 0)``.  (The IIW judgements themselves are not available offline.)"""
    rng = np.random.default_rng(seed)
    blob = np.full((n, max_comparisons + 1, 1, 6), np.nan)
    for i in range(n):
        count = int(rng.integers(min_count, max_comparisons + 1))
        blob[i, :count, 0, 0:4] = rng.uniform(0.0, 1.0, (count, 4)) * (1.0 - 1e-9)
        blob[i, :count, 0, 4] = rng.integers(0, 3, count)
        blob[i, :count, 0, 5] = rng.uniform(0.05, 3.0, count)
        blob[i, max_comparisons, 0, 0:3] = (count, float(100000 + i), 0.0)
    return blob
