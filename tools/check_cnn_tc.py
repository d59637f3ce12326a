#!/usr/bin/env python
"""Quick correctness probe of the tcgen05 CNN kernel against the CPU oracle (and the FP32 kernel)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from reflectance_filtering_b200 import cnn, synth  # noqa: E402

net = cnn.default_net()
print("mode:", "FP32 kernel" if os.environ.get("RF_CNN_FP32") == "1" else "tcgen05 kernel")
for (h, w, kind) in [(4, 4, "stress"), (16, 8, "stress"), (37, 53, "stress"), (384, 512, "natural"), (384, 512, "stress")]:
    img = synth.GENERATORS[kind](h, w, 7)
    r = cnn.get_reflectance_caffe(net, img)
    ref = oracle.mlp_forward(net.mlp, img)
    r64 = oracle.mlp_forward_f64(net.mlp, img)
    print(h, w, kind, "max|gpu-oracle32| %.3e  max|gpu-f64| %.3e  nan %d" % (np.abs(r - ref).max(),
          np.abs(r - r64).max(), int(np.isnan(r).sum())))
imgs = torch.from_numpy(synth.batch("natural", 4, 96, 80, 2)).cuda()
f32, u8 = net.forward_device(imgs, want_f32=True, want_u8=True)
print("u8 == trunc(f32*255):", bool(torch.equal(u8, (f32 * 255).to(torch.uint8))))
