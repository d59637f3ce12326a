import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import oracle
from reflectance_filtering_b200 import filters, synth
for (h, w, r, sc) in [(33, 47, 7, 3), (33, 47, 7, 1), (40, 56, 7, 3), (96, 120, 45, 1), (64, 100, 20, 3)]:
    img = synth.natural(h, w, 702); gd = synth.flat(h, w, 802)
    src = img if sc == 3 else img[:, :, 0].copy()
    ref = oracle.guided(gd, src, r, 3.0)
    for fill in (0.0, float("nan"), 3e38, -3e38):
        ws = torch.full((8 << 20,), fill, dtype=torch.float32, device="cuda").view(torch.uint8)
        tsrc = torch.from_numpy(src[None]).cuda()
        out = filters.guided_device(torch.from_numpy(gd[None]).cuda(), tsrc, r, 3.0, workspace=ws).cpu().numpy()[0]
        d = np.abs(out.astype(int) - ref.astype(int))
        print((h, w, r, sc), "fill", fill, "max diff", d.max(), "bad", int((d > 0).sum()), np.argwhere(d > 1)[:3].tolist())
