#!/usr/bin/env python
"""Randomised shapes through the fast guided-filter path (TMA rings, row plans) against the oracle.
usage: python tools/gf_stress.py [cases] [seed]   -- run it under `compute-sanitizer --tool memcheck` for the record."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from reflectance_filtering_b200 import filters, synth  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 120
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
worst = (0, 0.0)
for k in range(cases):
    r = int(rng.integers(8, 65))
    h = int(rng.integers(r + 1, 4 * r + 40))
    w = int(rng.integers(max(52, r + 4), 900))
    sc = int(rng.choice([1, 3]))
    n = int(rng.integers(1, 4))
    iters = int(rng.choice([1, 1, 2, 3]))
    eps = float(rng.choice([0.5, 3.0, 7.0, 50.0]))
    gd = np.stack([synth.flat(h, w, 5000 + 7 * k + i) for i in range(n)])
    src = np.stack([synth.natural(h, w, 6000 + 7 * k + i) for i in range(n)])
    src = src if sc == 3 else np.ascontiguousarray(src[..., k % 3])
    out = filters.guided_device(torch.from_numpy(gd).cuda(), torch.from_numpy(src).cuda(), r, eps, iterations=iters).cpu().numpy()
    ref = src[n - 1]
    for _ in range(iters):
        ref = oracle.guided(gd[n - 1], ref, r, eps)
    d = np.abs(out[n - 1].astype(int) - ref.astype(int))
    frac = float((d > 0).mean())
    worst = (max(worst[0], int(d.max())), max(worst[1], frac))
    if d.max() > 1 or frac > 5e-3:
        print("FAIL case", k, dict(r=r, h=h, w=w, sc=sc, n=n, iters=iters, eps=eps), "max", int(d.max()), "frac", frac)
        sys.exit(1)
print("gf_stress: %d cases ok, worst max %d LSB, worst fraction of differing bytes %.2e" % (cases, worst[0], worst[1]))
