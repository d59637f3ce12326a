#!/usr/bin/env python
"""One gray joint-bilateral call (64 x 512x384, c20 s22; RF_BF_PROBE_SIGMAS=c,s overrides), for ncu captures while tuning."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reflectance_filtering_b200 import filters, synth  # noqa: E402
n, h, w = 64, 384, 512
sc, ss = [float(v) for v in os.environ.get("RF_BF_PROBE_SIGMAS", "20,22").split(",")]
base = np.stack([synth.natural(h, w, 30 + i)[..., 1] for i in range(4)])
gray = torch.from_numpy(np.ascontiguousarray(np.stack([base[i % 4] for i in range(n)]))).cuda()
out = torch.empty_like(gray)
for _ in range(3):
    filters.joint_bilateral_device(gray, gray, sc, ss, out=out, gray_replicated=True)
torch.cuda.synchronize()
