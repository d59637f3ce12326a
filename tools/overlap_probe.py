#!/usr/bin/env python
"""Does CNN(k+1) overlap BF(k)?  K steps of 64 x 512x384 CNN -> BF(CNN,CNN): one stream vs CNN on a second
(high-priority) stream one step ahead.  VERDICT r1 item 6."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reflectance_filtering_b200 import cnn, filters, pipeline, synth  # noqa: E402

n, h, w, K = 64, 384, 512, 12
net = cnn.default_net(0)
pipe = pipeline.Pipeline(net)
base = np.stack([synth.natural(h, w, 2000 + i) for i in range(8)])
pool = [torch.from_numpy(np.stack([np.roll(base[(i + p) % 8], 5 * p + i, axis=1) for i in range(n)])).cuda() for p in range(6)]
outs = [torch.empty((n, h, w), dtype=torch.uint8, device="cuda") for _ in range(6)]
r8s = [torch.empty((n, h, w), dtype=torch.uint8, device="cuda") for _ in range(3)]


def serial():
    for k in range(K):
        pipe.cnn_bf(pool[k % 6], 20.0, 22.0, out=outs[k % 6], scratch=r8s[0])


def pipelined(prio):
    main = torch.cuda.current_stream()
    side = torch.cuda.Stream(priority=-1 if prio else 0)
    side.wait_stream(main)
    ready = [torch.cuda.Event() for _ in range(K)]
    done = [torch.cuda.Event() for _ in range(K)]
    for k in range(K):
        with torch.cuda.stream(side):
            if k >= 3:
                side.wait_event(done[k - 3])     # the scratch plane of step k-3 is free again
            pipe.reflectance_u8(pool[k % 6], out=r8s[k % 3])
            ready[k].record(side)
        main.wait_event(ready[k])
        filters.joint_bilateral_device(r8s[k % 3], r8s[k % 3], 20.0, 22.0, gray_replicated=True, out=outs[k % 6])
        done[k].record(main)
    main.wait_stream(side)


def timeit(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K


serial()
ref = [o.clone() for o in outs]
print("serial            %.3f ms/step" % timeit(serial))
for prio in (False, True):
    ms = timeit(lambda: pipelined(prio))
    same = all(torch.equal(a, b) for a, b in zip(ref, outs))
    print("pipelined prio=%d  %.3f ms/step  identical=%s" % (prio, ms, same))
