#!/usr/bin/env python
"""Per-step timings of Pipeline.run_host (pinned host in/out) for a few chunk / stream settings."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reflectance_filtering_b200 import cnn, pipeline, synth  # noqa: E402

H, W, B = 384, 512, 64
pipe = pipeline.Pipeline(cnn.default_net())
base = np.stack([synth.natural(H, W, 2000 + i) for i in range(8)])
pool = torch.empty((7, B, H, W, 3), dtype=torch.uint8, pin_memory=True)
for p in range(7):
    for i in range(B):
        pool[p, i] = torch.from_numpy(np.roll(base[(i + p) % 8], 7 * p + i, axis=1))
out = torch.empty((7, B, H, W), dtype=torch.uint8, pin_memory=True)
dev_in = pool[0].cuda()
torch.cuda.synchronize()

# raw copy bandwidth
for name, fn in [("h2d 37.7MB", lambda: pool[1].cuda(non_blocking=True)),
                 ("d2h 12.6MB", lambda: out[0].copy_(dev_in[..., 0], non_blocking=True))]:
    ts = []
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print(name, ["%.2f" % t for t in ts], "ms")

for chunk, ns in [(16, 3), (16, 2), (32, 2), (64, 1), (8, 4)]:
    ts = []
    for i in range(10):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pipe.run_host("cnn_bf", pool[i % 7], out[i % 7], chunk=chunk, n_streams=ns, sigma_color=20.0, sigma_spatial=22.0)
        ts.append((time.perf_counter() - t0) * 1e3)
    print("chunk", chunk, "streams", ns, ["%.1f" % t for t in ts], "ms/step")
# device-resident reference
ts = []
d_out = torch.empty((B, H, W), dtype=torch.uint8, device="cuda")
for i in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pipe.cnn_bf(dev_in, 20.0, 22.0, out=d_out)
    torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
print("device-resident", ["%.1f" % t for t in ts])
print("cpu count", os.cpu_count(), "load", os.getloadavg())
