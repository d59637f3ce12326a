#!/usr/bin/env python
"""Runs the five BASELINE.json configurations at their full sizes and prints one JSON line each.

  python tools/run_configs.py                        # 1 GPU
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/run_configs.py   # N GPUs

Images of a configuration are sharded in contiguous chunks over the ranks (no communication on the pixel path);
throughput = all images of the configuration / max-over-ranks CUDA-event time.  Inputs are device resident
(a small set of seeded synthetic images, rolled to make every image of the batch distinct); every configuration
ends with a parity spot check of one image against the CPU oracle where the oracle finishes in seconds, and with
size-independent properties otherwise.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402  (checker only)
from reflectance_filtering_b200 import cnn, filters, pipeline, synth  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)


def make_batch(gen, n, h, w, seed, distinct=8):
    base = torch.from_numpy(np.stack([gen(h, w, seed + i) for i in range(min(n, distinct))])).to(dev)
    out = torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev)
    for i in range(n):
        out[i] = torch.roll(base[i % base.shape[0]], shifts=(5 * (i // base.shape[0])) % w, dims=1)
    return out


def timed(fn, reps=3):
    fn()
    best = None
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        best = ms if best is None else min(best, ms)
    return best


def lsb(a, b):
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    return int(d.max()), float((d != 0).mean())


def emit(**kw):
    if rank == 0:
        print(json.dumps(kw), flush=True)


net = cnn.default_net(dev)
pipe = pipeline.Pipeline(net)

# ---- config 1: one 512x384 colour image, BF c20 s22 with a copy of itself as joint ----------------------
img = torch.from_numpy(synth.natural(384, 512, 1000)[None]).to(dev)
joint = img.clone()
out = torch.empty_like(img)
ms = timed(lambda: filters.joint_bilateral_device(joint, img, 20.0, 22.0, out=out), reps=5)
ref = oracle.joint_bilateral(joint[0].cpu().numpy(), img[0].cpu().numpy(), -1, 20.0, 22.0)
mx, frac = lsb(out[0].cpu().numpy(), ref)
emit(config=1, workload="1 x 512x384 colour, BF c20 s22, joint = copy", ms=ms, MPps=0.196608 / (ms * 1e-3),
     parity={"max_lsb": mx, "frac_differing": frac, "vs": "oracle == cv2.bilateralFilter"})

# ---- config 2: one 512x384 image, CNN -> BF(CNN,CNN) -------------------------------------------------------
res = None


def cfg2():
    global res
    res = pipe.cnn_bf(img, 20.0, 22.0)


ms = timed(cfg2, reps=5)
r8 = pipe.reflectance_u8(img)[0].cpu().numpy()
g3 = np.repeat(r8[:, :, None], 3, axis=2)
mx, frac = lsb(res[0].cpu().numpy(), oracle.joint_bilateral(g3.copy(), g3, -1, 20.0, 22.0)[:, :, 0])
emit(config=2, workload="1 x 512x384, CNN -> BF(CNN,CNN) c20 s22", ms=ms, MPps=0.196608 / (ms * 1e-3),
     parity={"max_lsb": mx, "frac_differing": frac, "vs": "oracle on the GPU's own CNN bytes"})

# ---- config 3: 64 x 512x384, CNN -> GF(CNN, flat) c3 s45 x3 --------------------------------------------------
n3 = 64
lo, hi = pipeline.shard_range(n3, rank, world)
if hi > lo:
    imgs = make_batch(synth.natural, hi - lo, 384, 512, 3000 + lo)
    guides = make_batch(synth.flat, hi - lo, 384, 512, 3500 + lo)
res3 = None


def cfg3():
    global res3
    if hi > lo:
        res3 = pipe.cnn_gf(imgs, guides, 3.0, 45.0, iterations=3)


ms = timed(cfg3)
par = None
if rank == 0:
    cur = pipe.reflectance_u8(imgs[:1])[0].cpu().numpy()
    g0 = guides[0].cpu().numpy()
    for _ in range(3):
        cur = oracle.guided(g0, cur, 45, 3.0)
    mx, frac = lsb(res3[0].cpu().numpy(), cur)
    par = {"max_lsb": mx, "frac_differing": frac, "vs": "oracle x3 with uint8 re-quantisation (image 0)"}
emit(config=3, workload="64 x 512x384, CNN -> GF(CNN, flat) c3 s45 x3", n_gpus=world, ms=ms,
     MPps=n3 * 0.196608 / (ms * 1e-3), parity=par)
del imgs, guides, res3

# ---- config 4: 5,230 x 1024x768, CNN -> BF c20 s22, contiguous shards -----------------------------------------
n4 = int(os.environ.get("RF_CFG4_IMAGES", "5230"))
lo, hi = pipeline.shard_range(n4, rank, world)
chunk = 512  # images resident at a time per rank (1.2 GB); the loop is the shard's "folder"
ims = make_batch(synth.natural, min(chunk, hi - lo), 768, 1024, 4000 + lo, distinct=4)
outb = torch.empty(ims.shape[:3], dtype=torch.uint8, device=dev)
scr = torch.empty_like(outb)


def cfg4():
    done = 0
    while done < hi - lo:
        m = min(chunk, hi - lo - done)
        pipe.cnn_bf(ims[:m], 20.0, 22.0, out=outb[:m], scratch=scr[:m])
        done += m


ms = timed(cfg4, reps=2)
# properties at full size: batch == per-image, constant rows stay put is covered in tests; check determinism
a = outb[:2].clone()
pipe.cnn_bf(ims[:2], 20.0, 22.0, out=outb[:2], scratch=scr[:2])
emit(config=4, workload="%d x 1024x768, CNN -> BF(CNN,CNN) c20 s22, sharded" % n4, n_gpus=world, ms=ms,
     MPps=n4 * 0.786432 / (ms * 1e-3), parity={"rerun_identical": bool(torch.equal(a, outb[:2]))})
del ims, outb, scr

# ---- config 5: 256 x 3840x2160, CNN -> BF c15 s28 and CNN -> GF c3 s45 -------------------------------------------
n5 = int(os.environ.get("RF_CFG5_IMAGES", "256"))
lo, hi = pipeline.shard_range(n5, rank, world)
chunk = 32
ims = make_batch(synth.natural, min(chunk, hi - lo), 2160, 3840, 5000 + lo, distinct=2)
gds = make_batch(synth.flat, min(chunk, hi - lo), 2160, 3840, 5500 + lo, distinct=2)
outb = torch.empty(ims.shape[:3], dtype=torch.uint8, device=dev)
scr = torch.empty_like(outb)


def cfg5_bf():
    done = 0
    while done < hi - lo:
        m = min(chunk, hi - lo - done)
        pipe.cnn_bf(ims[:m], 15.0, 28.0, out=outb[:m], scratch=scr[:m])
        done += m


def cfg5_gf():
    done = 0
    while done < hi - lo:
        m = min(chunk, hi - lo - done)
        pipe.cnn_gf(ims[:m], gds[:m], 3.0, 45.0, iterations=1)
        done += m


ms_bf = timed(cfg5_bf, reps=2)
ms_gf = timed(cfg5_gf, reps=2)
# parity on a 256x320 crop-sized problem is covered by the tests; here: a 4K constant image is a fixed point of
# both filters, and the BF output of a gray plane stays within the input range
const = torch.full((1, 2160, 3840), 137, dtype=torch.uint8, device=dev)
fix_bf = bool(torch.equal(filters.joint_bilateral_device(const, const, 15.0, 28.0, gray_replicated=True), const))
fix_gf = bool(torch.equal(filters.guided_device(gds[:1], const, 45, 3.0), const))
emit(config=5, workload="%d x 3840x2160, CNN -> BF c15 s28 | CNN -> GF c3 s45" % n5, n_gpus=world,
     bf={"ms": ms_bf, "MPps": n5 * 8.2944 / (ms_bf * 1e-3)}, gf={"ms": ms_gf, "MPps": n5 * 8.2944 / (ms_gf * 1e-3)},
     parity={"constant_fixed_point_bf": fix_bf, "constant_fixed_point_gf": fix_gf})

if world > 1:
    dist.destroy_process_group()
