#!/usr/bin/env python
"""Benchmark of the CNN -> joint bilateral filter hot path (BASELINE.json metric: megapixels/s).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
  python bench.py --impl reference [--steps K] [--warmup W]      # host-CPU reference-equivalent path

Workload (BASELINE.json configs[1]): synthetic 512x384 sRGB images -> CNN reflectance ->
trunc(r*255) -> BF(CNN, CNN) c20 s22.  One step = one batch of `--batch` such images per GPU
(weak scaling: every rank processes its own shard of the batch, no communication on the pixel
path).  Inputs rotate through a pool larger than L2 so no step re-reads cached inputs.

`value`   : device-resident uint8 in -> uint8 out, CUDA-event time over exactly K steps, max over ranks.
`e2e`     : the same steps through Pipeline.run_host: pinned HOST buffers in, HOST buffers out, H2D
            and D2H copies inside the timed region.
`roofline`: the dominant kernel (bf_gray2_kernel), timed per launch with CUDA events in a second pass
            over the same steps.
`cpu_baseline`: the reference-equivalent CPU path (cv2.dnn on the real prototxt/caffemodel +
            cv2.bilateralFilter, bit-identical to the restated jointBilateralFilter_8u) on a
            bounded sample, all host threads.
Secondary objects of the same JSON line (rank 0; skipped with --no-extra): `config3_cnn_gf_x3` (device resident and
through host buffers), `config4_strong` (a FIXED global batch of 1024x768 images split over the ranks: strong
scaling), `config5_slice` (4K images: BF c15 s28 and GF c3 s45), `single_image` (configs[0]/[1] latency),
`roofline_cnn`, `roofline_bf_color`, `parity`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W = 384, 512
SIGMA_COLOR, SIGMA_SPATIAL = 20.0, 22.0
CONFIG_ID = 2
METRIC = "megapixels/sec end-to-end CNN->BF(CNN,CNN) c20 s22 (device-resident uint8 in -> uint8 out)"
L2_BYTES = 126 * 1024 * 1024


def make_config(batch, world):
    """The same dict from both arms (the driver compares them); everything arm-specific is a top-level key."""
    return {"workload": "configs[1]: 512x384 CNN -> trunc u8 -> BF(CNN,CNN) c20 s22 (r=33, 3409 taps/px)",
            "images_per_step_per_gpu": batch, "height": H, "width": W, "sigma_color": SIGMA_COLOR,
            "sigma_spatial": SIGMA_SPATIAL,
            "parallelism": "dp%d, contiguous image shards, no collective" % world}


# What the parity claims of this repo rest on (oracle/README.md): printed with every record so
# that no reader takes "+-1 LSB against our restatement" for "+-1 LSB against OpenCV-contrib".
PARITY = {
    "cnn": "max abs error < 1e-5 vs cv2.dnn on the reference prototxt + caffemodel (tolerance 1e-3)",
    "bf": "+-1 LSB vs a restatement that is bit-equal to cv2.bilateralFilter (joint == src) on every test image",
    "gf": "unpinned vs ximgproc: +-1 LSB vs two independent in-house restatements of guided_filter.cpp and within "
          "1 LSB of a float64 He-et-al. formulation; no cv2.ximgproc build exists offline to pin against",
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# clock sampling during the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler(object):
    """SM clock and throttle reasons of one GPU, sampled DURING the timed region: an NVML polling thread
    (every 10 ms, so even a 100 ms region gets samples), or a `nvidia-smi -lms` child if NVML is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []
        self.sm, self.mx, self.reasons = [], [], set()
        self.handle = None
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                import torch
                uuid = "GPU-" + str(torch.cuda.get_device_properties(device).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                idx = device.index if hasattr(device, "index") else int(device)
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
        except Exception:
            self.handle = None

    def sample_now(self):
        if self.handle is None:
            return
        nv = self.nv
        try:
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
            if not self.mx:   # a constant of the board: asked once, the polling loop stays fast
                self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)))
            mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            for name, bit in (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown),
                              ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                              ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                              ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap)):
                if mask & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _poll(self):
        while not self._stop.is_set():
            self.sample_now()
            self._stop.wait(0.01)

    def start(self):
        if self.handle is not None:
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        try:
            idx = self.device.index if hasattr(self.device, "index") else int(self.device)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.handle is not None:
            self._stop.set()
            self.t.join(timeout=2)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None,
                    "sm_max_mhz": max(self.mx) if self.mx else None, "samples": len(self.sm),
                    "reasons": sorted(self.reasons), "source": "nvml, 10 ms polling during the timed region"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 100"}


# ---------------------------------------------------------------------------------------------
# reference-equivalent CPU path (also the --impl reference arm)
# ---------------------------------------------------------------------------------------------
def cpu_reference_step(images, dnn):
    """One image at a time, as the reference CLI chain does: CNN (cv2.dnn on the real artefacts,
    reference input transform), trunc(r*255), replicate to 3 channels (cv2.imread of the gray PNG),
    bilateral c20 s22 with a copy of itself as joint (== cv2.bilateralFilter)."""
    from oracle import anchors
    outs = []
    for img in images:
        r = dnn.forward(img)
        g3 = np.repeat(anchors.quantize_like_imwrite(r)[:, :, None], 3, axis=2)
        outs.append(anchors.bilateral_self(g3, SIGMA_COLOR, SIGMA_SPATIAL))
    return outs


def time_cpu(n_images: int, steps: int, warmup: int):
    import cv2
    from oracle import anchors
    from reflectance_filtering_b200 import synth
    from reflectance_filtering_b200.caffe_model import DEFAULT_CAFFEMODEL, DEFAULT_PROTOTXT
    cores = os.cpu_count() or 1
    cv2.setNumThreads(cores)
    dnn = anchors.DnnNet(DEFAULT_PROTOTXT, DEFAULT_CAFFEMODEL)
    imgs = [synth.natural(H, W, 1000 * CONFIG_ID + i) for i in range(n_images)]
    for _ in range(warmup):
        cpu_reference_step(imgs[:1], dnn)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_step(imgs, dnn)
    dt = time.perf_counter() - t0
    mp = steps * n_images * H * W / 1e6
    return mp / dt, dt / steps, cores, cv2.getNumThreads()


def run_reference(args, rank: int):
    if rank != 0:
        return
    n_img = 1
    mps, s_per_step, cores, cvt = time_cpu(n_img, args.steps, min(args.warmup, 2))
    line = {
        "impl": "reference", "metric": METRIC, "value": mps, "unit": "MP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(args.batch, max(1, args.gpus)),
        "reference_sample": "each step times %d image(s) of this workload on the host CPU (bounded sample)" % n_img,
        "parity": PARITY,
        "cpu_baseline": {"value": mps, "unit": "MP/s", "cores": cores, "kind": "port",
                         "sample": "%d step(s) x %d image(s) of 512x384: cv2.dnn forward on the reference "
                                   "prototxt+caffemodel, trunc, cv2.bilateralFilter c20 s22; cv2 threads=%d"
                                   % (args.steps, n_img, cvt)},
        "e2e": {"value": mps, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------
# secondary measurements of the CUDA arm (the other BASELINE.json configs, at bench-sized slices)
# ---------------------------------------------------------------------------------------------
# Per-launch DRAM traffic and executed instructions of the kernels below come from committed `ncu --set full`
# captures, not from this run: every such number carries the file it was read from.
NCU = {
    "bf_gray2": {"dram_bytes": 12630784, "source": "profiles/r02_bf_gray2_ncu_full.txt (64 x 512x384)"},
    # the six launches of a 3-iteration call at 64 x 512x384 (pass_a<full+stats>, pass_b, 2 x (pass_a<source only>,
    # pass_b)) + pack: 2.23 GB read + 1.18 GB written, 500.9 M warp instructions
    "gf_x3": {"dram_bytes": 3412000000, "warp_instructions": 500.9e6, "source": "profiles/r02_gf_final_ncu_full.txt"},
}


def timed(fn, reps, barrier, max_over_ranks):
    """Mean device time of `fn` over `reps` back-to-back calls (CUDA events, max over ranks), after one warm-up."""
    import torch
    fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    barrier()
    return max_over_ranks(e0.elapsed_time(e1)) / reps


def rolled_pool(gen, n, h, w, seed, device, distinct=8):
    """`n` images from `distinct` seeded generator calls, decorrelated by rolls (generation is host-bound)."""
    import torch
    base = [gen(h, w, seed + i) for i in range(min(n, distinct))]
    arr = np.stack([np.roll(base[i % len(base)], shift=11 * (i // len(base)), axis=1) for i in range(n)])
    return torch.from_numpy(arr).to(device)


def measure_extras(args, rank, world, device, pipe, dev_pool, host_pool, barrier, max_over_ranks):
    import torch
    from reflectance_filtering_b200 import filters, pipeline, synth
    B = args.batch
    reps = max(3, args.steps // 4)
    x = {}

    # ---- configs[0] shape: 3-channel BF with a copy of the image as joint (the colour kernel), batch B ----------
    jcopy = dev_pool[0].clone()
    cout = torch.empty_like(jcopy)
    x["bf_color_ms"] = timed(lambda: filters.joint_bilateral_device(jcopy, dev_pool[0], SIGMA_COLOR, SIGMA_SPATIAL, out=cout),
                             reps, barrier, max_over_ranks)
    x["bf_color_batch"] = B
    del jcopy, cout

    # ---- configs[2]: CNN -> GF(CNN, flat) c3 s45 x3, batch of 64 ------------------------------------------------
    GB = args.gf_batch
    d_flat = rolled_pool(synth.flat, GB, H, W, 1000 * 3 + 500 + rank * GB, device, distinct=16)
    d_imgs = dev_pool[0][:GB] if GB <= B else dev_pool.reshape(-1, H, W, 3)[:GB]
    x["gf_batch"] = GB
    x["cnn_gf_x3_ms"] = timed(lambda: pipe.cnn_gf(d_imgs, d_flat, 3.0, 45.0, iterations=3), reps, barrier, max_over_ranks)
    r8 = pipe.reflectance_u8(d_imgs)
    tmp = torch.empty_like(r8)
    x["gf_single_ms"] = timed(lambda: filters.guided_device(d_flat, r8, 45, 3.0, out=tmp), reps, barrier, max_over_ranks)
    x["gf_x3_ms"] = timed(lambda: filters.guided_device(d_flat, r8, 45, 3.0, out=tmp, iterations=3), reps, barrier,
                          max_over_ranks)
    # the same through pinned host buffers (images AND guides go up, the gray result comes down)
    h_img = host_pool[0][:GB] if GB <= B else host_pool.reshape(-1, H, W, 3)[:GB]
    h_gd = torch.empty((GB, H, W, 3), dtype=torch.uint8, pin_memory=True)
    h_gd.copy_(d_flat)
    h_out = torch.empty((GB, H, W), dtype=torch.uint8, pin_memory=True)
    t_wall = [0.0]

    def gf_host():
        t0 = time.perf_counter()
        pipe.run_host("cnn_gf", h_img, h_out, guides=h_gd, chunk=args.chunk, n_streams=4, sigma_color=3.0,
                      sigma_spatial=45.0, iterations=3)
        t_wall[0] += time.perf_counter() - t0

    for _ in range(3):   # the first calls allocate the per-stream buffers and scratch
        gf_host()
    t_wall[0] = 0.0
    e_reps = max(reps, 5)
    ev_ms = timed(gf_host, e_reps, barrier, max_over_ranks)
    x["cnn_gf_x3_e2e_ms"] = max(ev_ms, max_over_ranks(t_wall[0] * 1e3 / (e_reps + 1)))
    x["cnn_gf_x3_e2e_ok"] = bool(torch.equal(h_out.to(device), pipe.cnn_gf(d_imgs, d_flat, 3.0, 45.0, iterations=3)))
    del d_flat, r8, tmp

    # ---- configs[3] slice, STRONG scaling: a fixed global batch of 1024x768 images, contiguous shards -------------
    G4, H4, W4 = args.cfg4_images, 768, 1024
    lo, hi = pipeline.shard_range(G4, rank, world)
    imgs4 = rolled_pool(synth.natural, hi - lo, H4, W4, 1000 * 4 + lo, device)
    out4 = torch.empty((hi - lo, H4, W4), dtype=torch.uint8, device=device)
    r84 = torch.empty_like(out4)
    x["cfg4"] = {"global_images": G4, "per_rank": hi - lo,
                 "ms": timed(lambda: pipe.cnn_bf(imgs4, SIGMA_COLOR, SIGMA_SPATIAL, out=out4, scratch=r84), 3, barrier,
                             max_over_ranks)}
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    pipe.reflectance_u8(imgs4, out=r84)
    ev[1].record()
    filters.joint_bilateral_device(r84, r84, SIGMA_COLOR, SIGMA_SPATIAL, gray_replicated=True, out=out4)
    ev[2].record()
    barrier()
    x["cfg4"]["cnn_ms"], x["cfg4"]["bf_ms"] = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    del imgs4, out4, r84

    # ---- configs[4] slice: 4K images, CNN -> BF c15 s28 and CNN -> GF c3 s45 (weak: args.cfg5_images per rank) ----
    N5, H5, W5 = args.cfg5_images, 2160, 3840
    imgs5 = rolled_pool(synth.natural, N5, H5, W5, 1000 * 5 + rank * N5, device, distinct=2)
    flat5 = rolled_pool(synth.flat, N5, H5, W5, 1000 * 5 + 500 + rank * N5, device, distinct=2)
    out5 = torch.empty((N5, H5, W5), dtype=torch.uint8, device=device)
    r85 = torch.empty_like(out5)
    c5 = {"images_per_rank": N5}
    c5["cnn_bf_ms"] = timed(lambda: pipe.cnn_bf(imgs5, 15.0, 28.0, out=out5, scratch=r85), 2, barrier, max_over_ranks)
    c5["bf_ms"] = timed(lambda: filters.joint_bilateral_device(r85, r85, 15.0, 28.0, gray_replicated=True, out=out5), 2,
                        barrier, max_over_ranks)
    c5["cnn_gf_ms"] = timed(lambda: pipe.cnn_gf(imgs5, flat5, 3.0, 45.0, iterations=1), 3, barrier, max_over_ranks)
    c5["gf_ms"] = timed(lambda: filters.guided_device(flat5, r85, 45, 3.0, out=out5), 3, barrier, max_over_ranks)
    x["cfg5"] = c5
    del imgs5, flat5, out5, r85

    # ---- single-image latency: configs[0] (colour BF, joint = copy) and configs[1] (CNN -> BF(CNN,CNN)) ---------------
    one = dev_pool[0][:1].contiguous()
    jone = one.clone()
    o3 = torch.empty_like(one)
    o1 = torch.empty((1, H, W), dtype=torch.uint8, device=device)
    s1 = torch.empty_like(o1)
    x["single"] = {
        "config0_bf_color_ms": timed(lambda: filters.joint_bilateral_device(jone, one, SIGMA_COLOR, SIGMA_SPATIAL, out=o3),
                                     10, barrier, max_over_ranks),
        "config1_cnn_bf_ms": timed(lambda: pipe.cnn_bf(one, SIGMA_COLOR, SIGMA_SPATIAL, out=o1, scratch=s1), 10, barrier,
                                   max_over_ranks)}
    return x


def format_extras(x, world, peaks, peak_src, sms, f_max, taps):
    """JSON objects of the secondary measurements (rank 0)."""
    import ctypes as C
    from reflectance_filtering_b200 import _native
    out = {}
    sfu_peak = sms * 16 * f_max
    alu_peak = sms * 128 * f_max
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    px = x["bf_color_batch"] * H * W
    ctaps = float(px) * taps / (x["bf_color_ms"] * 1e-3)
    out["roofline_bf_color"] = {
        "kernel": "bf_color_kernel (configs[0] shape: 3-channel joint = copy of the 3-channel source), batch %d"
                  % x["bf_color_batch"],
        "bound": "fp32", "launch_ms": x["bf_color_ms"], "achieved": ctaps * 10 / 1e9, "peak": alu_peak / 1e9,
        "unit": "G FP32 lane-ops/s (SURVEY 8d: 10 lane-ops + 1 exp per tap)", "frac": ctaps * 10 / alu_peak,
        "sfu": {"achieved_Gtap_s": ctaps / 1e9, "peak_Gtap_s": sfu_peak / 1e9, "frac": ctaps / sfu_peak},
        "traffic": None}

    # configs[2].  Algorithmic bytes per pixel per iteration, 1-channel src (SURVEY 8d counts 44 with
    # a 3-byte source read in both passes): pass A reads 3 (guide) + 1 (src) and writes 16 (a0,a1,a2,b); pass B reads
    # 16 + 3 (guide) and writes 1  => 40 B/px
    gpx = x["gf_batch"] * H * W
    bpp = 40.0
    ach = gpx * bpp * 3 / (x["gf_x3_ms"] * 1e-3) / 1e9
    t_hbm_ms = gpx * bpp * 3 / (hbm * 1e9) * 1e3
    roof = {"kernel": "rf_guided_iterated_u8, 3 iterations: gf2 pack + pass_a<full+stats> + pass_b, then "
                      "2 x (pass_a<source only> + pass_b)", "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s",
            "frac": ach / hbm, "bytes_per_pixel_per_iteration": bpp, "launch_ms": x["gf_x3_ms"],
            "single_iteration_call_ms": x["gf_single_ms"], "traffic": NCU["gf_x3"]["dram_bytes"],
            "traffic_source": NCU["gf_x3"]["source"], "peak_source": peak_src}
    if NCU["gf_x3"]["warp_instructions"] and x["gf_batch"] == 64:
        # what also bounds the call: the executed warp instructions of its seven launches (ncu) at one instruction per
        # scheduler and clock; the call cannot be faster than the slower of the two bounds
        t_issue_ms = NCU["gf_x3"]["warp_instructions"] / (sms * 4 * f_max) * 1e3
        roof["issue_bound"] = {"warp_instructions": NCU["gf_x3"]["warp_instructions"],
                               "thread_instructions_per_pixel_per_iteration":
                                   NCU["gf_x3"]["warp_instructions"] * 32.0 / (gpx * 3),
                               "ms_at_full_issue_rate": t_issue_ms, "ms_at_hbm_peak": t_hbm_ms,
                               "frac_vs_max_of_both": max(t_issue_ms, t_hbm_ms) / x["gf_x3_ms"],
                               "source": NCU["gf_x3"]["source"]}
    h2d, d2h = world * gpx * 6, world * gpx
    out["config3_cnn_gf_x3"] = {
        "workload": "configs[2]: %d x 512x384, CNN -> GF(CNN, flat guide) c3.0 s45.0 (r=45), 3 iterations, "
                    "uint8 re-quantisation between iterations" % x["gf_batch"],
        "value": world * gpx / (x["cnn_gf_x3_ms"] * 1e-3) / 1e6, "unit": "MP/s", "ms_per_step": x["cnn_gf_x3_ms"],
        "e2e": {"value": world * gpx / (x["cnn_gf_x3_e2e_ms"] * 1e-3) / 1e6, "unit": "MP/s",
                "ms_per_step": x["cnn_gf_x3_e2e_ms"], "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "host_link_GBs": (h2d + d2h) / world / (x["cnn_gf_x3_e2e_ms"] * 1e-3) / 1e9,
                "api": "Pipeline.run_host('cnn_gf', pinned images + pinned guides -> pinned uint8[N,H,W])",
                "matches_device_path": x["cnn_gf_x3_e2e_ok"],
                "note": "7 bytes cross the host link per pixel (55 GB/s per direction on this box, tools/h2d_probe.py): the "
                        "copies are as long as the kernels and only partly overlap them"},
        "roofline": roof}

    r_, t_ = C.c_int(), C.c_int()
    c4 = x["cfg4"]
    px4 = c4["global_images"] * 768 * 1024
    out["config4_strong"] = {
        "workload": "configs[3] slice: a FIXED global batch of %d images of 1024x768, CNN -> BF(CNN,CNN) c20 s22, contiguous "
                    "shards over the ranks, no collective" % c4["global_images"],
        "scaling": "strong", "value": px4 / (c4["ms"] * 1e-3) / 1e6, "unit": "MP/s", "ms": c4["ms"],
        "images_on_rank0": c4["per_rank"],
        "rank0_kernels_ms": {"mlp_tc_kernel": c4["cnn_ms"], "bf_gray2_kernel": c4["bf_ms"]},
        "roofline": {"kernel": "bf_gray2_kernel", "bound": "sfu",
                     "frac": c4["per_rank"] * 768 * 1024 * float(taps) / (c4["bf_ms"] * 1e-3) / sfu_peak}}

    _native.lib().rf_joint_bilateral_geometry(28.0, -1, C.byref(r_), C.byref(t_))
    c5 = x["cfg5"]
    px5 = c5["images_per_rank"] * 2160 * 3840
    out["config5_slice"] = {
        "workload": "configs[4] slice: %d image(s) of 3840x2160 per GPU" % c5["images_per_rank"], "scaling": "weak",
        "cnn_bf_c15_s28": {"value": world * px5 / (c5["cnn_bf_ms"] * 1e-3) / 1e6, "unit": "MP/s", "ms": c5["cnn_bf_ms"],
                           "taps_per_pixel": t_.value,
                           "roofline": {"kernel": "bf_gray2_kernel", "bound": "sfu", "launch_ms": c5["bf_ms"],
                                        "frac": px5 * float(t_.value) / (c5["bf_ms"] * 1e-3) / sfu_peak}},
        "cnn_gf_c3_s45": {"value": world * px5 / (c5["cnn_gf_ms"] * 1e-3) / 1e6, "unit": "MP/s", "ms": c5["cnn_gf_ms"],
                          "roofline": {"kernel": "rf_guided_u8 (gf2 pack + pass_a + pass_b)", "bound": "hbm",
                                       "launch_ms": c5["gf_ms"], "achieved": px5 * 40.0 / (c5["gf_ms"] * 1e-3) / 1e9,
                                       "peak": hbm, "unit": "GB/s", "frac": px5 * 40.0 / (c5["gf_ms"] * 1e-3) / 1e9 / hbm}}}
    out["single_image"] = dict(x["single"], note="one 512x384 image per call, device resident, mean of 10 back-to-back calls")
    return out


# ---------------------------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------------------------
def run_cuda(args, rank: int, local_rank: int, world: int):
    import torch
    import torch.distributed as dist
    from reflectance_filtering_b200 import _native, cnn, filters, pipeline, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)

    B = args.batch
    net = cnn.default_net(device)
    pipe = pipeline.Pipeline(net)

    # ---- synthetic pool, larger than L2, built from per-image seeded generators ------------------
    n_distinct = min(B, 32)
    lo = rank * B  # every rank draws its own shard of the (world * B)-image batch
    base = np.stack([synth.natural(H, W, 1000 * CONFIG_ID + lo + i) for i in range(n_distinct)])
    batch_bytes = B * H * W * 3
    n_pool = max(2, int(np.ceil(2.0 * L2_BYTES / batch_bytes)))
    host_pool = torch.empty((n_pool, B, H, W, 3), dtype=torch.uint8, pin_memory=True)
    for p in range(n_pool):
        for i in range(B):
            src = base[(i + p) % n_distinct]
            host_pool[p, i] = torch.from_numpy(np.roll(src, shift=(7 * p + 3 * (i // n_distinct)) % W, axis=1))
    dev_pool = host_pool.to(device)
    out_pool = torch.empty((n_pool, B, H, W), dtype=torch.uint8, device=device)
    host_out = torch.empty((n_pool, B, H, W), dtype=torch.uint8, pin_memory=True)
    torch.cuda.synchronize()

    def step(i):
        p = i % n_pool
        pipe.cnn_bf(dev_pool[p], SIGMA_COLOR, SIGMA_SPATIAL, out=out_pool[p])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing -------------------------------------------------------------------
    for i in range(args.warmup):
        step(i)
    sampler = ClockSampler(device)
    barrier()
    sampler.start()
    launches0 = _native.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i)
        marks[i].record()  # per-step marks for best / median; the headline uses e0..e1 over all K steps
    e1.record()
    sampler.sample_now()  # kernels of the timed region are still in flight here
    barrier()
    launches = _native.launch_count() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop()
    per_step = [(e0 if i == 0 else marks[i - 1]).elapsed_time(marks[i]) for i in range(args.steps)]
    px_step = B * H * W
    value = world * px_step * args.steps / (ms_total * 1e-3) / 1e6

    # ---- end to end through host buffers ----------------------------------------------------------
    def e2e_step(i):
        p = i % n_pool
        pipe.run_host("cnn_bf", host_pool[p], host_out[p], chunk=args.chunk, n_streams=4,
                      sigma_color=SIGMA_COLOR, sigma_spatial=SIGMA_SPATIAL)

    for i in range(min(args.warmup, 3)):
        e2e_step(i)
    barrier()
    e0.record()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(args.warmup + i)
    e1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), wall_ms))
    e2e_value = world * px_step * args.steps / (e2e_ms * 1e-3) / 1e6
    e2e_ok = bool(torch.equal(host_out[(args.warmup + args.steps - 1) % n_pool].to(device),
                              out_pool[(args.warmup + args.steps - 1) % n_pool]))

    # ---- per-kernel pass: the same steps with events around each launch ------------------------------
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        p = (args.warmup + i) % n_pool
        ev[i][0].record()
        r8 = pipe.reflectance_u8(dev_pool[p])
        ev[i][1].record()
        filters.joint_bilateral_device(r8, r8, SIGMA_COLOR, SIGMA_SPATIAL, d=-1, gray_replicated=True,
                                       out=out_pool[p])
        ev[i][2].record()
    barrier()
    cnn_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    bf_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))

    extra = None
    if not args.no_extra:
        extra = measure_extras(args, rank, world, device, pipe, dev_pool, host_pool, barrier, max_over_ranks)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = measured_peaks()
    import ctypes as C
    r_, t_ = C.c_int(), C.c_int()
    _native.lib().rf_joint_bilateral_geometry(SIGMA_SPATIAL, -1, C.byref(r_), C.byref(t_))
    taps = t_.value
    sms = torch.cuda.get_device_properties(device).multi_processor_count
    f_max = float(peaks.get("sm_max_mhz", 1965.0)) * 1e6
    f_obs = (clocks.get("sm_mhz") or 0) * 1e6
    taps_per_launch = float(px_step) * taps
    # gray fast path: 4 FP32 lane-ops + 1 exp per tap; MUFU.EX2 issues 16 lanes/clk/SM
    sfu_peak = sms * 16 * f_max
    alu_peak = sms * 128 * f_max
    achieved_taps = taps_per_launch / (bf_ms * 1e-3)
    roofline = {
        "kernel": "bf_gray2_kernel", "bound": "sfu", "achieved": achieved_taps / 1e9, "peak": sfu_peak / 1e9,
        "unit": "Gtap/s (1 MUFU.EX2 per tap; peak = SMs*16*sm_max_mhz)", "frac": achieved_taps / sfu_peak,
        "note": "the roofline assumes one MUFU.EX2 per tap; the kernel evaluates one weight pair in eight with an FMA-pipe "
                "polynomial instead, so the XU pipe itself is ~90 % busy (DESIGN.md K3)",
        "frac_at_observed_clock": (achieved_taps / (sms * 16 * f_obs)) if f_obs else None,
        "fp32_lane_ops": {"per_tap": 4, "achieved_Gop_s": achieved_taps * 4 / 1e9, "peak_Gop_s": alu_peak / 1e9,
                          "frac": achieved_taps * 4 / alu_peak},
        "taps_per_pixel": taps, "launch_ms": bf_ms, "share_of_step": bf_ms / (bf_ms + cnn_ms),
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch at this batch size, from a committed `ncu --set
        # full` capture (12.63 MB read, 0 B written: the output is still in L2 when the kernel ends; algorithmic
        # bytes are 2 * px = 25.2 MB) -- not measured in this run
        "traffic": NCU["bf_gray2"]["dram_bytes"] if B == 64 else None,
        "traffic_source": NCU["bf_gray2"]["source"],
        "peak_source": "SM count x unit width x %s clock" % peak_src,
    }
    cnn_flops = 8704.0 * px_step
    tf32_peak = float(peaks.get("bf16_tflops", 1590.0)) / 2.0  # dense TF32 = half the measured bf16 rate
    cnn_roof = {"kernel": "mlp_tc_kernel (tcgen05.mma kind::tf32, A from TMEM, 3x split operands)", "bound": "tensor",
                "launch_ms": cnn_ms, "achieved": cnn_flops / (cnn_ms * 1e-3) / 1e12, "peak": tf32_peak,
                "unit": "TFLOP/s (algorithmic 8,704 FLOP/px; 3xTF32 split operands: 59 MMAs of 128x32x8 per 128-pixel tile, "
                        "conv0 and all biases included)",
                "frac": cnn_flops / (cnn_ms * 1e-3) / 1e12 / tf32_peak,
                "vs_fp32_cuda_core_peak": cnn_flops / (cnn_ms * 1e-3) / (alu_peak * 2),
                "note": "latency bound by construction: K = N = 32 per layer, activations round-trip TMEM -> registers "
                        "(ReLU, fuse dot, hi/lo split) -> TMEM between layers; four tile pipelines per SM (TMEM "
                        "capacity) overlap each other (DESIGN.md K2)"}

    cpu = None
    if world == 1 and not args.no_cpu:
        mps, s_per, cores, cvt = time_cpu(args.cpu_images, 1, 1)
        cpu = {"value": mps, "unit": "MP/s", "cores": cores, "kind": "port",
               "sample": "%d image(s) of 512x384 through cv2.dnn (reference prototxt+caffemodel) + trunc + "
                         "cv2.bilateralFilter c20 s22, cv2 threads=%d, %.1f s" % (args.cpu_images, cvt, s_per)}

    line = {
        "metric": METRIC, "value": value, "unit": "MP/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(B, world),
        "cache": "inputs rotate through a %d-batch pool (%.0f MB > L2)" % (n_pool, n_pool * batch_bytes / 1e6),
        "parity": PARITY,
        "e2e": {"value": e2e_value, "unit": "MP/s", "h2d_bytes_per_step": world * batch_bytes,
                "d2h_bytes_per_step": world * px_step, "ms_per_step": e2e_ms / args.steps,
                "api": "Pipeline.run_host(pinned host uint8[N,H,W,3] -> pinned host uint8[N,H,W])",
                "matches_device_path": e2e_ok},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "roofline_cnn": cnn_roof,
        "cpu_baseline": cpu,
    }
    line["step_ms"] = {"best": float(np.min(per_step)), "median": float(np.median(per_step)),
                       "note": "per-step CUDA-event marks on rank 0 inside the timed region"}
    if extra is not None:
        line.update(format_extras(extra, world, peaks, peak_src, sms, f_max, taps))
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_RECORD_FD = None


def emit(line):
    """The one JSON line of the run, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _RECORD_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RECORD_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="images per step per GPU")
    ap.add_argument("--chunk", type=int, default=32, help="largest images-per-chunk of the e2e path (H2D / compute / D2H; the first chunks are smaller)")
    ap.add_argument("--cpu-images", type=int, default=6, help="sample size of the cpu_baseline leg")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", "--no-gf", dest="no_extra", action="store_true",
                    help="skip the secondary measurements (configs[0], [2], [3], [4] slices, single-image latency)")
    ap.add_argument("--gf-batch", type=int, default=64)
    ap.add_argument("--cfg4-images", type=int, default=128, help="global batch of the strong-scaling configs[3] slice")
    ap.add_argument("--cfg5-images", type=int, default=2, help="4K images per GPU of the configs[4] slice")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries exactly ONE line (the JSON record of rank 0).  Native libraries print there too (NCCL's
    # version banner at communicator creation), so file descriptor 1 is pointed at stderr for the run and the
    # record is written to the saved descriptor at the end.
    global _RECORD_FD
    sys.stdout.flush()
    _RECORD_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_cuda(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
