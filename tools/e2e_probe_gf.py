#!/usr/bin/env python
"""Per-call wall time of Pipeline.run_host('cnn_gf', ...) on 64 x 512x384 for several chunk sizes (development probe)."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reflectance_filtering_b200 import cnn, pipeline, synth  # noqa: E402
n, h, w = 64, 384, 512
pipe = pipeline.Pipeline(cnn.default_net(0))
base = np.stack([synth.natural(h, w, 100 + i) for i in range(8)])
flat = np.stack([synth.flat(h, w, 200 + i) for i in range(8)])
h_img = torch.from_numpy(np.stack([base[i % 8] for i in range(n)])).pin_memory()
h_gd = torch.from_numpy(np.stack([flat[i % 8] for i in range(n)])).pin_memory()
h_out = torch.empty((n, h, w), dtype=torch.uint8).pin_memory()
for chunk in [int(c) for c in (sys.argv[1:] or ["4", "8", "16", "32", "64"])]:
    ts = []
    for rep in range(8):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pipe.run_host("cnn_gf", h_img, h_out, guides=h_gd, chunk=chunk, n_streams=4, sigma_color=3.0, sigma_spatial=45.0,
                      iterations=3)
        ts.append((time.perf_counter() - t0) * 1e3)
    print("chunk %2d: per-call ms %s  -> best %.0f MP/s" % (chunk, " ".join("%.2f" % t for t in ts), n * h * w / min(ts) / 1e3))
