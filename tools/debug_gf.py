import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import oracle
from reflectance_filtering_b200 import filters, synth
img = synth.natural(33, 47, 702); gd = synth.flat(33, 47, 802)
img4 = synth.natural(33, 47, 704); gd4 = synth.flat(33, 47, 804)
ref = oracle.guided(gd, img, 7, 3.0)
a = filters.apply_filter("guided", img, gd, 3.0, 7.0)
print("apply_filter n=1 vs oracle:", np.abs(a.astype(int) - ref).max(), (a != ref).sum())
d1 = filters.guided_device(torch.from_numpy(gd[None]).cuda(), torch.from_numpy(img[None]).cuda(), 7, 3.0).cpu().numpy()[0]
print("device n=1 vs oracle:", np.abs(d1.astype(int) - ref).max(), (d1 != ref).sum())
g2 = torch.from_numpy(np.stack([gd, gd4])).cuda(); s2 = torch.from_numpy(np.stack([img, img4])).cuda()
d2 = filters.guided_device(g2, s2, 7, 3.0).cpu().numpy()
print("device n=2 img0 vs oracle:", np.abs(d2[0].astype(int) - ref).max(), (d2[0] != ref).sum())
print("device n=2 img1 vs oracle:", np.abs(d2[1].astype(int) - oracle.guided(gd4, img4, 7, 3.0)).max())
bad = np.argwhere(np.abs(a.astype(int) - ref) > 1)
print("bad positions (first 10):", bad[:10].tolist(), "count", len(bad))
cur = img
for it in range(3):
    cur = filters.apply_filter("guided", cur, gd, 3.0, 7.0)
    ref = oracle.guided(gd, ref if it else img, 7, 3.0) if it else ref
    print("iter", it, "max diff vs oracle chain:", np.abs(cur.astype(int) - ref).max())
print("---- poisoned allocator ----")
filters._ws_cache.clear()
for fill in (float("nan"), 1e30, -1e30):
    junk = [torch.full((1 << 20,), fill, dtype=torch.float32, device="cuda") for _ in range(64)]
    del junk
    filters._ws_cache.clear()
    ref = oracle.guided(gd, img, 7, 3.0)
    a = filters.apply_filter("guided", img, gd, 3.0, 7.0)
    bad = np.argwhere(np.abs(a.astype(int) - ref) > 0)
    print("fill", fill, "max diff", np.abs(a.astype(int) - ref).max(), "count", len(bad), bad[:6].tolist())
    g1 = filters.apply_filter("guided", img[:, :, 0].copy(), gd, 3.0, 7.0)
    r1 = oracle.guided(gd, img[:, :, 0].copy(), 7, 3.0)
    print("   gray: max diff", np.abs(g1.astype(int) - r1).max())
