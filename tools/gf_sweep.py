#!/usr/bin/env python
"""Row-segment sweep of the guided-filter fast path (development tool): one process per setting, because the
library reads RF_GF2_SEGS_A / RF_GF2_SEGS_B once.  usage: python tools/gf_sweep.py"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, json, numpy as np, torch
sys.path.insert(0, %r)
from reflectance_filtering_b200 import filters, synth
n, h, w = 64, 384, 512
base = np.stack([synth.flat(h, w, 10 + i) for i in range(4)])
guide = torch.from_numpy(np.stack([base[i %% 4] for i in range(n)])).cuda()
base = np.stack([synth.natural(h, w, 20 + i) for i in range(4)])
src = torch.from_numpy(np.stack([base[i %% 4] for i in range(n)])[..., 0].copy()).cuda()
out = torch.empty_like(src)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(fn):
    for _ in range(2): fn()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)
print(json.dumps({"single_ms": t(lambda: filters.guided_device(guide, src, 45, 3.0, out=out)),
                  "x3_ms": t(lambda: filters.guided_device(guide, src, 45, 3.0, out=out, iterations=3))}))
''' % ROOT
for sa, sb in [(0, 0), (4, 1), (4, 2), (4, 3), (4, 4), (2, 6), (3, 6), (6, 6), (3, 3), (2, 2)]:
    env = dict(os.environ)
    if sa: env["RF_GF2_SEGS_A"] = str(sa)
    if sb: env["RF_GF2_SEGS_B"] = str(sb)
    r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True, timeout=300)
    print("segs_a=%d segs_b=%d" % (sa, sb), r.stdout.strip() or r.stderr[-300:], flush=True)
