#!/usr/bin/env python
"""Randomised shapes / radii / sigmas through the tiled joint-bilateral kernels (gray, gray-replicated, colour, distinct
joint) against the oracle.  usage: python tools/bf_stress.py [cases] [seed]"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from reflectance_filtering_b200 import filters, synth  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 80
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 5)
worst = (0, 0.0)
for k in range(cases):
    ss = float(rng.uniform(0.7, 30.0))           # radius 1 .. 45
    sc_ = float(rng.choice([3.0, 8.0, 20.0, 60.0]))
    h = int(rng.integers(1, 140))
    w = int(rng.integers(1, 200))
    jc = int(rng.choice([1, 3]))
    sc = int(rng.choice([1, 3]))
    sep = bool(rng.integers(0, 2)) or jc != sc
    gray_rep = (jc == 1 and sc == 1 and not sep and bool(rng.integers(0, 2)))
    gen = synth.stress if k % 3 == 0 else synth.natural
    joint = gen(h, w, 7000 + k)
    src = gen(h, w, 8000 + k) if sep else joint
    joint = joint if jc == 3 else np.ascontiguousarray(joint[:, :, 1])
    src = (src if sc == 3 else np.ascontiguousarray(src[:, :, 1])) if sep else joint
    dj = torch.from_numpy(np.ascontiguousarray(joint))[None].cuda()
    ds = dj if not sep else torch.from_numpy(np.ascontiguousarray(src))[None].cuda()
    out = filters.joint_bilateral_device(dj, ds, sc_, ss, gray_replicated=gray_rep).cpu().numpy()[0]
    if gray_rep:
        j3 = np.repeat(joint[:, :, None], 3, axis=2)
        ref = oracle.joint_bilateral(j3, j3, -1, sc_, ss)[:, :, 0]
    else:
        ref = oracle.joint_bilateral(joint, src, -1, sc_, ss)
    d = np.abs(out.astype(int) - ref.reshape(out.shape).astype(int))
    frac = float((d > 0).mean())
    worst = (max(worst[0], int(d.max())), max(worst[1], frac if d.size > 2000 else 0.0))
    if d.max() > 1 or (d.size > 2000 and frac > 5e-3):
        print("FAIL case", k, dict(ss=ss, sc=sc_, h=h, w=w, jc=jc, sc_ch=sc, sep=sep, gray_rep=gray_rep), int(d.max()), frac)
        sys.exit(1)
print("bf_stress: %d cases ok, worst max %d LSB, worst fraction of differing bytes (images > 2000 values) %.2e" % (cases, worst[0], worst[1]))
