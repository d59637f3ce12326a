#!/usr/bin/env python
"""Per-kernel timings on the current GPU (CUDA events, L2 flushed between launches).

  python tools/kbench.py gf | bf | cnn | all
Prints one JSON object per case; used while tuning, not a contract benchmark (that is bench.py).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reflectance_filtering_b200 import cnn, filters, synth  # noqa: E402

PEAK_HBM = 6545.0
try:
    PEAK_HBM = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timeit(fn, reps=6, warm=2):
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def tile_batch(gen, n, h, w, seed):
    base = np.stack([gen(h, w, seed + i) for i in range(min(n, 4))])
    return torch.from_numpy(np.stack([base[i % len(base)] for i in range(n)])).cuda()


def bench_gf():
    for (n, h, w) in [(64, 384, 512), (8, 2160, 3840), (1, 384, 512)]:
        guide = tile_batch(synth.flat, n, h, w, 10)
        src3 = tile_batch(synth.natural, n, h, w, 20)
        src1 = src3[..., 0].contiguous()
        for name, src, bpp in [("gray", src1, 40.0), ("color", src3, 108.0)]:
            out = torch.empty_like(src)
            med, mn = timeit(lambda: filters.guided_device(guide, src, 45, 3.0, out=out))
            px = n * h * w
            print(json.dumps({"kernel": "gf r=45 " + name, "shape": [n, h, w], "ms": med, "ms_min": mn,
                              "GB/s": px * bpp / (med * 1e-3) / 1e9, "frac_hbm": px * bpp / (med * 1e-3) / 1e9 / PEAK_HBM,
                              "Mpx/s": px / (med * 1e-3) / 1e6}))
            med3, mn3 = timeit(lambda: filters.guided_device(guide, src, 45, 3.0, out=out, iterations=3))
            print(json.dumps({"kernel": "gf r=45 " + name + " x3 (guide statistics cached)", "shape": [n, h, w],
                              "ms": med3, "ms_min": mn3, "ms_per_iteration": med3 / 3,
                              "frac_hbm": px * bpp * 3 / (med3 * 1e-3) / 1e9 / PEAK_HBM}))


def bench_bf():
    sfu_peak = 148 * 16 * 1.965e9
    for (n, h, w) in [(64, 384, 512), (1, 384, 512), (4, 2160, 3840)]:
        img = tile_batch(synth.natural, n, h, w, 30)
        gray = img[..., 1].contiguous()
        jcopy = img.clone()
        for name, fn, taps in [
            ("gray c20 s22", lambda o: filters.joint_bilateral_device(gray, gray, 20, 22, gray_replicated=True, out=o), 3409),
            ("gray c15 s28", lambda o: filters.joint_bilateral_device(gray, gray, 15, 28, gray_replicated=True, out=o), 5525),
            ("color self c20 s22", lambda o: filters.joint_bilateral_device(img, img, 20, 22, out=o), 3409),
            ("color joint-copy c20 s22", lambda o: filters.joint_bilateral_device(jcopy, img, 20, 22, out=o), 3409),
        ]:
            out = torch.empty_like(gray if name.startswith("gray") else img)
            med, mn = timeit(lambda: fn(out), reps=4, warm=1)
            px = n * h * w
            print(json.dumps({"kernel": "bf " + name, "shape": [n, h, w], "ms": med, "ms_min": mn,
                              "Gtap/s": px * taps / (med * 1e-3) / 1e9, "frac_sfu": px * taps / (med * 1e-3) / sfu_peak,
                              "Mpx/s": px / (med * 1e-3) / 1e6}))


def bench_cnn():
    net = cnn.default_net()
    for (n, h, w) in [(64, 384, 512), (1, 384, 512), (8, 2160, 3840)]:
        img = tile_batch(synth.natural, n, h, w, 40)
        med, mn = timeit(lambda: net.forward_device(img, want_f32=False, want_u8=True))
        px = n * h * w
        print(json.dumps({"kernel": "cnn u8 out", "shape": [n, h, w], "ms": med, "ms_min": mn,
                          "TFLOP/s": px * 8704 / (med * 1e-3) / 1e12, "Mpx/s": px / (med * 1e-3) / 1e6}))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("gf", "all"):
        bench_gf()
    if what in ("bf", "all"):
        bench_bf()
    if what in ("cnn", "all"):
        bench_cnn()
