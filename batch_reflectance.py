#!/usr/bin/env python
"""Batch front-end of the B200 reflectance-filtering path: a folder or glob of images in, the reference's
per-image output files out (``<name>-r.png``, ``<name>-r_<type>_c<sc>s<ss>.png``).

  python batch_reflectance.py --inputs photos/ --path_out out/ --filter_type=bilateral --sigma_color=20 --sigma_spatial=22
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 batch_reflectance.py ...   # one shard per GPU

An additive tool: the reference has no batch CLI on its inference path (its folder modes live in the training
helper).  Flags shared with ``filter_reflectance.py`` keep their names and meaning.
"""
from __future__ import print_function

import argparse
import json
import os
import time

from reflectance_filtering_b200 import batch


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--inputs", required=True, help="directory, glob pattern or single image")
    ap.add_argument("--path_out", required=True, help="existing output directory")
    ap.add_argument("--mode", default="decompose+filter", choices=["decompose", "filter", "decompose+filter"])
    ap.add_argument("--filter_type", default="bilateral", help="bilateral or guided")
    ap.add_argument("--sigma_color", type=float, default=20.0)
    ap.add_argument("--sigma_spatial", type=float, default=22.0)
    ap.add_argument("--guidance_dir", default=None, help="directory of guidance images with the inputs' base names "
                                                        "(default: every image guides itself)")
    ap.add_argument("--iterations", type=int, default=1, help="re-apply the filter to its own uint8 output")
    ap.add_argument("--chunk", type=int, default=16)
    ap.add_argument("--io_threads", type=int, default=8)
    ap.add_argument("--skip_existing", action="store_true")
    ap.add_argument("--colorized", action="store_true", help="also write <name>-r_colorized.png / <name>-s_colorized.png")
    args = ap.parse_args(argv)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    files = batch.list_inputs(args.inputs)
    t0 = time.perf_counter()
    res = batch.run_batch(files, args.path_out, mode=args.mode, filter_type=args.filter_type,
                          sigma_color=args.sigma_color, sigma_spatial=args.sigma_spatial, guidance=args.guidance_dir,
                          iterations=args.iterations, device=local, io_threads=args.io_threads, chunk=args.chunk,
                          rank=rank, world=world, skip_existing=args.skip_existing, colorized=args.colorized)
    dt = time.perf_counter() - t0
    print(json.dumps({"rank": rank, "world": world, "images": res["images"], "written": len(res["written"]),
                      "errors": res["errors"], "seconds": dt,
                      "MP_per_s_including_png_io": res["pixels"] / 1e6 / dt if dt > 0 else None}))
    return 0 if not res["errors"] else 1


if __name__ == "__main__":
    raise SystemExit(main())
