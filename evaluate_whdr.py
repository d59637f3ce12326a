#!/usr/bin/env python
"""Mean WHDR of the CNN reflectance before and after filtering, on one or several GPUs.

  python evaluate_whdr.py                                             # synthetic images + synthetic judgements
  python evaluate_whdr.py --images photos/ --comparisons blob.npy     # blob [n, max+1, 1, 6], sorted-file order
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 evaluate_whdr.py ...

The metric and blob layout are the reference's (training/layers/whdr_layer.py:253-287,
training/createNumpyArrayWithComparisonsForIIW.py:616-649); the IIW judgements themselves are not shipped, so the
default run draws synthetic comparisons (reflectance_filtering_b200.synth.comparisons).  Each rank evaluates a
contiguous shard of the images; the only communication is one all-reduce of [sum of WHDRs, image count] per line.
"""
from __future__ import print_function

import argparse
import json
import os

import cv2
import numpy as np
import torch

from reflectance_filtering_b200 import batch, cnn, device as dev, filters, pipeline, synth, whdr


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--images", default=None, help="directory or glob of equally sized images (default: synthetic)")
    ap.add_argument("--comparisons", default=None, help=".npy comparison blob [n, max+1, 1, 6]")
    ap.add_argument("--n", type=int, default=32, help="number of synthetic images")
    ap.add_argument("--height", type=int, default=384)
    ap.add_argument("--width", type=int, default=512)
    ap.add_argument("--delta", type=float, default=0.1)
    ap.add_argument("--sigma_color", type=float, default=20.0)
    ap.add_argument("--sigma_spatial", type=float, default=22.0)
    ap.add_argument("--chunk", type=int, default=16)
    args = ap.parse_args(argv)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    d = dev.bind_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1 and not torch.distributed.is_initialized():
        torch.distributed.init_process_group("nccl", device_id=d)

    files = batch.list_inputs(args.images) if args.images else None
    n = len(files) if files else args.n
    lo, hi = pipeline.shard_range(n, rank, world)
    if files:
        imgs = np.stack([cv2.imread(f) for f in files[lo:hi]]) if hi > lo else np.zeros((0, 1, 1, 3), np.uint8)
    else:
        imgs = synth.batch("natural", hi - lo, args.height, args.width, 6, start=lo)
    blob = np.load(args.comparisons) if args.comparisons else synth.comparisons(n, 1181, seed=6, min_count=100)
    if blob.shape[0] != n:
        raise ValueError("%d comparison blobs for %d images" % (blob.shape[0], n))
    blob = torch.from_numpy(np.ascontiguousarray(blob[lo:hi])).to(d)

    net = cnn.default_net(d)
    per = {"cnn_float32": [], "cnn_uint8": [], "cnn_bilateral": [], "cnn_guided": []}
    for s in range(0, hi - lo, args.chunk):
        x = torch.from_numpy(imgs[s:s + args.chunk]).to(d)
        cmp = blob[s:s + args.chunk]
        f32, u8 = net.forward_device(x, want_f32=True, want_u8=True)
        bf = filters.joint_bilateral_device(u8, u8, args.sigma_color, args.sigma_spatial, gray_replicated=True)
        gf = filters.guided_device(filters.replicate_gray_device(u8), u8, int(args.sigma_spatial), args.sigma_color)
        per["cnn_float32"].append(whdr.whdr_device(f32, cmp, args.delta))
        for key, t in (("cnn_uint8", u8), ("cnn_bilateral", bf), ("cnn_guided", gf)):
            per[key].append(whdr.whdr_device(t.to(torch.float32) / 255.0, cmp, args.delta))
    out = {"images": n, "world": world, "delta": args.delta,
           "comparisons": "file" if args.comparisons else "synthetic"}
    for key, parts in per.items():
        vals = torch.cat(parts) if parts else torch.zeros(0, dtype=torch.float64, device=d)
        out[key], count = whdr.reduce_mean(vals)
        assert count == n, (count, n)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
