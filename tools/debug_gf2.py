import os, sys, tempfile, pathlib
import numpy as np, torch, cv2
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import oracle
from reflectance_filtering_b200 import filters, synth, batch
td = pathlib.Path(tempfile.mkdtemp())
src_dir, out_a, gdir = td / "in", td / "batch", td / "guide"
for d in (src_dir, out_a, gdir): d.mkdir()
shapes = [(40, 56), (40, 56), (33, 47), (40, 56), (33, 47)]
for i, (h, w) in enumerate(shapes):
    cv2.imwrite(str(src_dir / ("im%d.png" % i)), synth.natural(h, w, 700 + i))
    cv2.imwrite(str(gdir / ("im%d.png" % i)), synth.flat(h, w, 800 + i))
files = batch.list_inputs(str(src_dir))
def chain(tag):
    img = cv2.imread(str(src_dir / "im2.png")); gd = cv2.imread(str(gdir / "im2.png"))
    cur = img; ref = img
    for it in range(3):
        cur = filters.apply_filter("guided", cur, gd, 3.0, 7.0)
        ref = oracle.guided(gd, ref, 7, 3.0)
        d = np.abs(cur.astype(int) - ref.astype(int))
        print(tag, "iter", it, "max diff", d.max(), "count>1", int((d > 1).sum()), np.argwhere(d > 1)[:4].tolist())
chain("before batch")
res = batch.run_batch(files, str(out_a), mode="filter", filter_type="guided", sigma_color=3, sigma_spatial=7,
                      guidance=str(gdir), iterations=3, chunk=4)
print("batch errors", res["errors"])
chain("after batch")
chain("after batch again")
