#!/usr/bin/env python
"""One guided-filter call per case, for `ncu --metrics gpu__time_duration.sum` launch lists while tuning."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reflectance_filtering_b200 import filters, synth  # noqa: E402
import numpy as np
n, h, w = 64, 384, 512
base = np.stack([synth.flat(h, w, 10 + i) for i in range(4)])
guide = torch.from_numpy(np.stack([base[i % 4] for i in range(n)])).cuda()
base = np.stack([synth.natural(h, w, 20 + i) for i in range(4)])
src3 = torch.from_numpy(np.stack([base[i % 4] for i in range(n)])).cuda()
src1 = src3[..., 0].contiguous()
for src in (src1, src3):
    out = torch.empty_like(src)
    for _ in range(2):
        filters.guided_device(guide, src, 45, 3.0, out=out)
    for _ in range(2):
        filters.guided_device(guide, src, 45, 3.0, out=out, iterations=3)
torch.cuda.synchronize()
