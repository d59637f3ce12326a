#!/usr/bin/env python
"""Pinned host <-> device copy bandwidth of this box (what bounds the CNN->GF e2e path: 7 bytes per pixel cross the link)."""
import torch
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    print("%s pinned 256 MiB: %.1f GB/s" % (name, 5 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
e1.record(); torch.cuda.synchronize()
print("both directions at once: %.1f GB/s each" % (5 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9))
