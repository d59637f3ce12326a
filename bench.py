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


def make_config(batch, world, taps=3409, extra=None):
    cfg = {"workload": "configs[1]: 512x384 CNN -> trunc u8 -> BF(CNN,CNN) c20 s22 (r=33, %d taps/px)" % taps,
           "images_per_step_per_gpu": batch, "height": H, "width": W, "sigma_color": SIGMA_COLOR,
           "sigma_spatial": SIGMA_SPATIAL,
           "parallelism": "dp%d, contiguous image shards, no collective" % world}
    if extra:
        cfg.update(extra)
    return cfg


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
        "config": make_config(args.batch, max(1, args.gpus),
                              extra={"reference_sample": "each step times %d image(s) of this workload on the host "
                                                         "CPU (bounded sample)" % n_img}),
        "cpu_baseline": {"value": mps, "unit": "MP/s", "cores": cores, "kind": "port",
                         "sample": "%d step(s) x %d image(s) of 512x384: cv2.dnn forward on the reference "
                                   "prototxt+caffemodel, trunc, cv2.bilateralFilter c20 s22; cv2 threads=%d"
                                   % (args.steps, n_img, cvt)},
        "e2e": {"value": mps, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


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

    # ---- configs[0] shape: 3-channel BF with a copy of the image as joint (the colour kernel) ---------
    bfc_ms = None
    if not args.no_gf:
        c_steps = max(3, args.steps // 4)
        jcopy = dev_pool[0].clone()
        cout = torch.empty_like(jcopy)
        filters.joint_bilateral_device(jcopy, dev_pool[0], SIGMA_COLOR, SIGMA_SPATIAL, out=cout)
        cev = [torch.cuda.Event(enable_timing=True) for _ in range(c_steps + 1)]
        barrier()
        cev[0].record()
        for i in range(c_steps):
            filters.joint_bilateral_device(jcopy, dev_pool[0], SIGMA_COLOR, SIGMA_SPATIAL, out=cout)
            cev[i + 1].record()
        barrier()
        bfc_ms = float(np.mean([cev[i].elapsed_time(cev[i + 1]) for i in range(c_steps)]))
        del jcopy, cout

    # ---- configs[2]: CNN -> GF(CNN, flat) c3 s45 x3, batch of 64 (secondary line, same JSON) ----------
    gf = None
    if not args.no_gf:
        GB = args.gf_batch
        flat = np.stack([synth.flat(H, W, 1000 * 3 + 500 + rank * GB + i) for i in range(min(GB, 16))])
        d_flat = torch.from_numpy(np.stack([flat[i % len(flat)] for i in range(GB)])).to(device)
        d_imgs = dev_pool[0][:GB] if GB <= B else dev_pool.reshape(-1, H, W, 3)[:GB]
        for _ in range(2):
            pipe.cnn_gf(d_imgs, d_flat, 3.0, 45.0, iterations=3)
        g_steps = max(3, args.steps // 4)
        barrier()
        e0.record()
        for _ in range(g_steps):
            pipe.cnn_gf(d_imgs, d_flat, 3.0, 45.0, iterations=3)
        e1.record()
        barrier()
        gf_ms = max_over_ranks(e0.elapsed_time(e1)) / g_steps
        r8 = pipe.reflectance_u8(d_imgs)
        gev = [torch.cuda.Event(enable_timing=True) for _ in range(g_steps + 1)]
        tmp = None
        barrier()
        gev[0].record()
        for i in range(g_steps):
            tmp = filters.guided_device(d_flat, r8, 45, 3.0, out=tmp)
            gev[i + 1].record()
        barrier()
        gf_iter_ms = float(np.mean([gev[i].elapsed_time(gev[i + 1]) for i in range(g_steps)]))
        # the three iterations as one call (rf_guided_iterated_u8: guide statistics computed once)
        barrier()
        gev[0].record()
        for i in range(g_steps):
            tmp = filters.guided_device(d_flat, r8, 45, 3.0, out=tmp, iterations=3)
            gev[i + 1].record()
        barrier()
        gf_x3_ms = float(np.mean([gev[i].elapsed_time(gev[i + 1]) for i in range(g_steps)]))
        gf = {"ms_per_step": gf_ms, "gf_iteration_ms": gf_iter_ms, "gf_x3_ms": gf_x3_ms, "batch": GB, "steps": g_steps}

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
        "frac_at_observed_clock": (achieved_taps / (sms * 16 * f_obs)) if f_obs else None,
        "fp32_lane_ops": {"per_tap": 4, "achieved_Gop_s": achieved_taps * 4 / 1e9, "peak_Gop_s": alu_peak / 1e9,
                          "frac": achieved_taps * 4 / alu_peak},
        "taps_per_pixel": taps, "launch_ms": bf_ms, "share_of_step": bf_ms / (bf_ms + cnn_ms),
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch at this batch size, from the committed
        # `ncu --set full` capture profiles/r01_bf_gray2_ncu_full.txt (12.63 MB read, 0 B written: the output is
        # still in L2 when the kernel ends); algorithmic bytes are 2 * px = 25.2 MB
        "traffic": 12633600 if B == 64 else None,
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
        "config": make_config(B, world, taps, extra={"cache": "inputs rotate through a %d-batch pool (%.0f MB > L2)"
                                                                 % (n_pool, n_pool * batch_bytes / 1e6)}),
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
    if bfc_ms is not None:
        ctaps = float(px_step) * taps / (bfc_ms * 1e-3)
        line["roofline_bf_color"] = {
            "kernel": "bf_color_kernel (configs[0] shape: 3-channel joint = copy of the 3-channel source), batch %d" % B,
            "bound": "fp32", "launch_ms": bfc_ms, "achieved": ctaps * 10 / 1e9, "peak": alu_peak / 1e9,
            "unit": "G FP32 lane-ops/s (SURVEY 8d: 10 lane-ops + 1 exp per tap)", "frac": ctaps * 10 / alu_peak,
            "sfu": {"achieved_Gtap_s": ctaps / 1e9, "peak_Gtap_s": sfu_peak / 1e9, "frac": ctaps / sfu_peak},
            "traffic": None}
    if gf is not None:
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        gpx = gf["batch"] * H * W
        # algorithmic bytes per pixel per iteration, 1-channel src: pass A reads 3 (guide)
        # + 1 (src) and writes 16 (a0,a1,a2,b); pass B reads 16 + 3 (guide) and writes 1  => 40 B/px
        bpp = 40.0
        ach = gpx * bpp * 3 / (gf["gf_x3_ms"] * 1e-3) / 1e9
        line["config3_cnn_gf_x3"] = {
            "workload": "configs[2]: %d x 512x384, CNN -> GF(CNN, flat guide) c3.0 s45.0 (r=45), 3 iterations, "
                        "uint8 re-quantisation between iterations" % gf["batch"],
            "value": world * gpx / (gf["ms_per_step"] * 1e-3) / 1e6, "unit": "MP/s", "ms_per_step": gf["ms_per_step"],
            "roofline": {"kernel": "rf_guided_iterated_u8, 3 iterations: gf2 pack + pass_a<full+stats> + pass_b, then "
                                   "2 x (pass_a<source only> + pass_b)", "bound": "hbm", "achieved": ach,
                         "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "bytes_per_pixel_per_iteration": bpp,
                         "launch_ms": gf["gf_x3_ms"], "single_iteration_call_ms": gf["gf_iteration_ms"],
                         # dram__bytes_read.sum + dram__bytes_write.sum of the seven launches at batch 64, from
                         # profiles/r01_gf_v8_ncu_full.txt: pass A<full+stats> 765 MB, pass B 856 MB, 2 x (pass A<source
                         # only> 807 MB + pass B 856 MB), pack ~0.11 GB (algorithmic: 3 x 40 B x 12.58 Mpx = 1.51 GB)
                         "traffic": 5060000000 if gf["batch"] == 64 else None, "peak_source": peak_src}}
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
    ap.add_argument("--chunk", type=int, default=8, help="images per H2D/compute/D2H chunk in the e2e path")
    ap.add_argument("--cpu-images", type=int, default=6, help="sample size of the cpu_baseline leg")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-gf", action="store_true", help="skip the secondary configs[2] (CNN->GF x3) measurement")
    ap.add_argument("--gf-batch", type=int, default=64)
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
