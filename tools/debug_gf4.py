import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import oracle
from reflectance_filtering_b200 import filters, synth
h, w, r, sc = 33, 47, 7, 1
img = synth.natural(h, w, 702); gd = synth.flat(h, w, 802)
src = img[:, :, 0].copy()
ws32 = torch.full((8 << 20,), float("nan"), dtype=torch.float32, device="cuda")
out = filters.guided_device(torch.from_numpy(gd[None]).cuda(), torch.from_numpy(src[None]).cuda(), r, 3.0, workspace=ws32.view(torch.uint8)).cpu().numpy()[0]
rh = 8; wp = (w + 2 * rh + 16 + 3) & ~3
plane = h * wp
ab = ws32[plane:plane + 4 * plane].cpu().numpy().reshape(4, h, wp)
print("wp", wp)
for k in range(4):
    nanc = np.isnan(ab[k]).any(axis=0)
    print("plane", k, "NaN columns:", np.nonzero(nanc)[0].tolist())
ref = oracle.guided(gd, src, r, 3.0)
d = np.abs(out.astype(int) - ref.astype(int))
print("bad cols", sorted(set(np.argwhere(d > 1)[:, 1].tolist())))
