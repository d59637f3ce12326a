"""Folder / glob front-end: PNG decode and encode on a host thread pool, overlapped with the GPU.

The reference processes one image per process launch; its folder modes live in the training helper
(/root/reference/training/train_with_barrista_helper.py:396-436, 711-1060: loop over files, swallow
per-file errors).  This module is the B200 equivalent for the inference chain:

    <name>.png --CNN--> <name>-r.png --filter--> <name>-r_<type>_c<sc>s<ss>.png

with exactly the file names and bytes the two reference CLIs would produce one image at a time
(decompose_with_trained_CNN.py:118, filter_reflectance.py:92-94).  Images are grouped by shape into
chunks, copied through pinned buffers on alternating CUDA streams, and a shard of the file list can be
given to each GPU / process (``rank``, ``world``) -- no communication between shards.
"""
from __future__ import annotations

import glob as _glob
import os
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Optional, Sequence

import cv2
import numpy as np
import torch

from . import cnn, device as dev, filters, pipeline


def list_inputs(spec: str | Sequence[str]) -> List[str]:
    """A directory (all *.png / *.jpg / *.jpeg / *.bmp in it), a glob pattern, or an explicit list."""
    if not isinstance(spec, str):
        return sorted(spec)
    if os.path.isdir(spec):
        out = []
        for ext in ("png", "jpg", "jpeg", "bmp", "PNG", "JPG"):
            out += _glob.glob(os.path.join(spec, "*." + ext))
        return sorted(set(out))
    return sorted(_glob.glob(spec))


def _stem(path: str) -> str:
    return os.path.splitext(os.path.basename(path))[0]


class BatchResult(dict):
    """``{"written": [...], "errors": {file: message}, "images": n, "pixels": n}``"""


def run_batch(files: Sequence[str], path_out: str, mode: str = "decompose+filter",
              filter_type: str = "bilateral", sigma_color: float = 20.0, sigma_spatial: float = 22.0,
              guidance: Optional[str] = None, iterations: int = 1, device=None, io_threads: int = 8,
              chunk: int = 16, rank: int = 0, world: int = 1, skip_existing: bool = False,
              colorized: bool = False) -> BatchResult:
    """Process ``files`` (this rank's contiguous shard of them).

    mode: ``decompose`` (CNN only, writes ``<name>-r.png``), ``filter`` (filter the given images with
    ``guidance``), ``decompose+filter`` (CNN, then filter the reflectance).  guidance: ``None`` = the image
    being filtered guides itself (BF(CNN,CNN)); otherwise a directory holding guidance images with the
    same base names as the inputs.  ``iterations`` > 1 re-applies the filter to its own uint8 output
    (the "3x GF" configuration); only the final result is written.  ``colorized``: in the decompose modes
    also write ``<name>-r_colorized.png`` and ``<name>-s_colorized.png`` as decompose_image does.
    """
    if mode not in ("decompose", "filter", "decompose+filter"):
        raise ValueError("mode must be 'decompose', 'filter' or 'decompose+filter'")
    if mode != "decompose":
        filters._validate(filter_type, sigma_color, sigma_spatial)
    if not os.path.isdir(path_out):
        raise Exception("Not able to write into {}, does the folder exist?".format(path_out))
    lo, hi = pipeline.shard_range(len(files), rank, world)
    files = list(files)[lo:hi]
    d = dev.bind_device(device)
    pipe = pipeline.Pipeline(cnn.default_net(d)) if mode != "filter" else None
    res = BatchResult(written=[], errors={}, images=0, pixels=0)
    suffix = "_{}_c{}s{}".format(filter_type, float(sigma_color), float(sigma_spatial))

    def out_names(f):
        stem = _stem(f)
        r_name = os.path.join(path_out, stem + "-r.png")
        if mode == "decompose":
            return r_name, None
        base = stem + "-r" if mode == "decompose+filter" else stem
        return r_name, os.path.join(path_out, base + suffix + ".png")

    def load(f):
        try:
            img = cv2.imread(f)
            if img is None:
                raise Exception("Input image not readable: {}".format(f))
            gd = None
            if guidance is not None and mode != "decompose":
                gpath = None
                for ext in (os.path.splitext(f)[1], ".png", ".jpg"):
                    cand = os.path.join(guidance, _stem(f) + ext)
                    if os.path.exists(cand):
                        gpath = cand
                        break
                gd = cv2.imread(gpath) if gpath else None
                if gd is None:
                    raise Exception("Input image not readable: {}".format(os.path.join(guidance, _stem(f) + ".*")))
                if gd.shape != img.shape:
                    raise ValueError("guidance and image sizes differ for {}".format(f))
            return f, img, gd, None
        except Exception as e:  # per-file isolation, as the reference's folder loop does
            return f, None, None, str(e)

    def save(path, arr):
        if not cv2.imwrite(path, arr):
            return path, "Not able to write {}, does the folder exist?".format(path)
        return path, None

    streams = [torch.cuda.Stream(device=d) for _ in range(2)]
    pinned: Dict[tuple, torch.Tensor] = {}

    def pin(tag, shape, slot):
        # At most two groups are in flight on the GPU, so three rotating buffers per (tag, shape) are never shared
        # by two groups' copies.  The ENCODERS read the output buffers later, on the save pool: process_group
        # waits for the saves of the group that used these buffers last (slot_saves) before it overwrites them.
        key = (tag, tuple(shape), slot % 3)
        buf = pinned.get(key)
        if buf is None:
            buf = pinned[key] = torch.empty(tuple(shape), dtype=torch.uint8, pin_memory=True)
        return buf

    pending_writes = []
    slot_saves: Dict[int, list] = {}  # slot % 3 -> save futures that still read that slot's pinned output buffers
    todo = []
    for f in files:
        r_name, f_name = out_names(f)
        final = f_name if f_name else r_name
        if skip_existing and os.path.exists(final):
            continue
        todo.append(f)

    def process_group(items, slot):
        """items: list of (file, img, guide) with one common shape."""
        n = len(items)
        h, w = items[0][1].shape[:2]
        s = streams[slot % 2]
        host = pin("in", (n, h, w, 3), slot)
        for i, (_, img, _) in enumerate(items):
            host[i] = torch.from_numpy(img)
        ghost = None
        if items[0][2] is not None:
            ghost = pin("guide", (n, h, w, 3), slot)
            for i, (_, _, gd) in enumerate(items):
                ghost[i] = torch.from_numpy(gd)
        with torch.cuda.stream(s):
            dimg = host.to(d, non_blocking=True)
            dgd = ghost.to(d, non_blocking=True) if ghost is not None else None
            outs = {}
            if mode == "filter":
                cur, gray = dimg, False
                # the per-file operator runs the single-channel kernels when the image (and, for the
                # bilateral filter, its guidance) is a gray image replicated to three channels; do the same
                flag = torch.ones(2, dtype=torch.int32, device=d)
                g1 = filters.extract_gray_device(dimg, flag[0:1])
                if filter_type == "bilateral" and dgd is not None:
                    g2 = filters.extract_gray_device(dgd, flag[1:2])
                ok = flag.tolist()
                if ok[0] and (filter_type == "guided" or dgd is None):
                    cur, gray = g1, True
                elif ok[0] and ok[1]:
                    # gray image guided by a different gray image: single-channel kernel, distinct joint
                    for _ in range(max(1, iterations)):
                        g1 = filters.joint_bilateral_device(g2, g1, sigma_color, sigma_spatial, gray_replicated=True)
                    outs["f"] = filters.replicate_gray_device(g1)
                    cur = None
            else:
                if colorized:
                    f32, cur = pipe.net.forward_device(dimg, want_f32=True, want_u8=True)
                    outs["rc"], outs["sc"] = cnn.colorize_device(dimg, f32)
                else:
                    cur = pipe.reflectance_u8(dimg)  # uint8 [n,h,w]: the bytes of <name>-r.png
                outs["r"] = cur
                gray = True
            if mode != "decompose" and cur is not None:
                fixed_guide = filter_type == "guided" and dgd is not None
                if fixed_guide:  # one guide for every iteration: its statistics are computed once
                    cur = filters.guided_device(dgd, cur, int(sigma_spatial), sigma_color, iterations=max(1, iterations))
                for _ in range(0 if fixed_guide else max(1, iterations)):
                    if filter_type == "bilateral":
                        if gray and dgd is None:
                            cur = filters.joint_bilateral_device(cur, cur, sigma_color, sigma_spatial, gray_replicated=True)
                        else:
                            src3 = filters.replicate_gray_device(cur) if gray else cur
                            jnt = dgd if dgd is not None else src3
                            cur, gray = filters.joint_bilateral_device(jnt, src3, sigma_color, sigma_spatial), False
                    else:
                        guide = filters.replicate_gray_device(cur) if gray else cur  # the image guides itself
                        cur = filters.guided_device(guide, cur, int(sigma_spatial), sigma_color)
                outs["f"] = filters.replicate_gray_device(cur) if gray else cur  # what cv2.imwrite gets
            for fut in slot_saves.pop(slot % 3, []):  # encoders of the group three slots back: done reading?
                fut.result()
            host_out = {k: pin("out_" + k, v.shape, slot) for k, v in outs.items()}
            for k, v in outs.items():
                host_out[k].copy_(v, non_blocking=True)
        return s, items, host_out, slot

    def flush(job, pool):
        s, items, host_out, slot = job
        s.synchronize()
        mine = slot_saves.setdefault(slot % 3, [])

        def submit(path, view):
            fut = pool.submit(save, path, view)
            pending_writes.append(fut)
            mine.append(fut)

        for i, (f, img, _) in enumerate(items):
            r_name, f_name = out_names(f)
            if "r" in host_out:
                submit(r_name, host_out["r"][i].numpy())
            for key, tail in (("rc", "-r_colorized.png"), ("sc", "-s_colorized.png")):
                if key in host_out:
                    submit(os.path.join(path_out, _stem(f) + tail), host_out[key][i].numpy())
            if "f" in host_out:
                submit(f_name, host_out["f"][i].numpy())
            res["images"] += 1
            res["pixels"] += img.shape[0] * img.shape[1]

    def load_ahead(pool, window):
        """Decoded files in order, with at most ``window`` decodes queued or waiting to be consumed (the whole
        folder is never held in RAM, and the decode queue cannot starve the encoders: they have their own pool)."""
        it = iter(todo)
        queue = []
        for f in it:
            queue.append(pool.submit(load, f))
            if len(queue) >= window:
                break
        while queue:
            out = queue.pop(0).result()
            for f in it:
                queue.append(pool.submit(load, f))
                break
            yield out

    n_io = max(1, io_threads)
    with ThreadPoolExecutor(max_workers=n_io) as load_pool, ThreadPoolExecutor(max_workers=n_io) as pool:
        loaded = load_ahead(load_pool, max(2 * chunk, 2 * n_io))
        groups: Dict[tuple, list] = {}
        in_flight = []
        slot = 0
        for f, img, gd, err in loaded:
            if err is not None:
                res["errors"][f] = err
                continue
            key = img.shape
            groups.setdefault(key, []).append((f, img, gd))
            if len(groups[key]) >= chunk:
                in_flight.append(process_group(groups.pop(key), slot))
                slot += 1
                if len(in_flight) > 1:
                    flush(in_flight.pop(0), pool)
        for key in list(groups):
            in_flight.append(process_group(groups.pop(key), slot))
            slot += 1
        for job in in_flight:
            flush(job, pool)
        for fut in pending_writes:
            path, err = fut.result()
            if err is None:
                res["written"].append(path)
            else:
                res["errors"][path] = err
    res["written"].sort()
    return res
