#!/usr/bin/env python
"""One colour joint-bilateral call (64 x 512x384, c20 s22, joint = copy of the source), for ncu captures."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reflectance_filtering_b200 import filters, synth  # noqa: E402
n, h, w = 64, 384, 512
base = np.stack([synth.natural(h, w, 30 + i) for i in range(4)])
img = torch.from_numpy(np.stack([base[i % 4] for i in range(n)])).cuda()
joint = img.clone()
out = torch.empty_like(img)
for _ in range(3):
    filters.joint_bilateral_device(joint, img, 20.0, 22.0, out=out)
torch.cuda.synchronize()
