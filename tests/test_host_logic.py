"""Host-side logic that runs without a GPU: model readers, image conventions against the
reference's own outputs, operator validation, CLI surface, C-ABI exports, sharding (gloo)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from reflectance_filtering_b200 import _native, caffe_model, image_utils as iu, pipeline, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "golden.npz"))


# ---- model artefacts --------------------------------------------------------------------------
def test_model_files_are_the_reference_artefacts():
    import hashlib
    h = hashlib.sha256(open(caffe_model.DEFAULT_CAFFEMODEL, "rb").read()).hexdigest()
    assert h == "51a7f4baabab139d560eff30c4bcd12c4ec74b0a4d5ee9884912a22cd7a90f76"


def test_pixel_mlp_structure(mlp):
    assert mlp.dims() == [3, 32, 32, 32, 32, 32]
    assert mlp.concat == [0, 1, 2, 3, 4]
    assert mlp.n_params == 4513 and mlp.macs_per_pixel == 4352
    assert mlp.layer_names == ["conv0", "conv1", "conv2", "conv3", "conv4", "fuse_skip_layers"]
    assert mlp.input_blob == "images" and mlp.output_blob == "reflectance_intensity"
    assert abs(mlp.fuse_b - 0.24792217) < 1e-7
    assert mlp.flat_params().shape == (4513,)
    # weight ranges of SURVEY B.1
    assert abs(mlp.hidden[0][0].min() + 2.608) < 1e-3 and abs(mlp.hidden[0][0].max() - 1.732) < 1e-3


def test_caffemodel_ignores_blobless_layers():
    blobs = caffe_model.read_caffemodel(caffe_model.DEFAULT_CAFFEMODEL)
    assert sorted(blobs) == ["conv0", "conv1", "conv2", "conv3", "conv4", "fuse_skip_layers"]
    assert blobs["conv0"][0].shape == (32, 3, 1, 1) and blobs["fuse_skip_layers"][0].shape == (1, 160, 1, 1)


def test_prototxt_rejects_unsupported_graphs(tmp_path):
    txt = open(caffe_model.DEFAULT_PROTOTXT).read()
    bad = tmp_path / "bad.prototxt"
    bad.write_text(txt.replace("kernel_size: 1", "kernel_size: 3", 1))
    with pytest.raises(ValueError, match="1x1"):
        caffe_model.build_pixel_mlp(str(bad))
    bad.write_text(txt.replace('type: "Sigmoid"', 'type: "TanH"'))
    with pytest.raises(ValueError, match="unsupported layer type"):
        caffe_model.build_pixel_mlp(str(bad))
    bad.write_text(txt.replace('name: "conv3"', 'name: "conv3_renamed"'))
    with pytest.raises(ValueError, match="no weights"):
        caffe_model.build_pixel_mlp(str(bad))


def test_prototxt_parser_handles_comments_and_strings():
    msg = caffe_model.parse_prototxt('a: 1 # c\nb { s: "x y" f: 1.5 e: TRAIN } b { s: "z" }')
    assert msg["a"] == [1] and msg["b"][0] == {"s": ["x y"], "f": [1.5], "e": ["TRAIN"]} and len(msg["b"]) == 2


# ---- image conventions against the reference's own outputs ------------------------------------
def test_srgb_helpers_bit_exact(G):
    x = G["srgb_in"]
    assert np.array_equal(iu.srgb_to_rgb(x), G["srgb_to_rgb"])
    assert np.array_equal(iu.rgb_to_srgb(x), G["rgb_to_srgb"])
    r32 = iu.srgb_to_rgb(x.astype(np.float32))
    assert r32.dtype == np.float32 and np.array_equal(r32, G["srgb_to_rgb_f32"])
    assert np.array_equal(iu.srgb_lut(), G["srgb_lut_f32"])
    # the quirk: the round trip is NOT the identity (max error 0.0324)
    assert 0.03 < np.abs(iu.rgb_to_srgb(iu.srgb_to_rgb(x)) - x).max() < 0.035


def test_imwrite_quantisation_and_png_round_trip(G, tmp_path):
    f = str(tmp_path / "g-r.png")
    iu.imwrite(f, G["imwrite_gray_in"])
    import cv2
    assert np.array_equal(cv2.imread(f, cv2.IMREAD_UNCHANGED), G["imwrite_gray_png"])
    assert np.array_equal(iu.imread(f), G["imread_gray_png"])
    u8 = G["imread_gray_png"]
    iu.imwrite(f, u8)
    assert np.array_equal(iu.imread(f), u8)  # uint8 is written untouched


def test_colorize_and_srgb_outputs(G, tmp_path):
    import cv2
    refl, shad = iu.colorize(G["imwrite_gray_in"], G["colorize_image"])
    assert np.array_equal(refl, G["colorize_reflectance"]) and np.array_equal(shad, G["colorize_shading"])
    fr, fs = str(tmp_path / "r.png"), str(tmp_path / "s.png")
    iu.imwrite(fr, refl, sRGB=True)
    iu.imwrite(fs, shad, sRGB=True)
    assert np.array_equal(cv2.imread(fr, cv2.IMREAD_UNCHANGED), G["colorize_r_png"])
    assert np.array_equal(cv2.imread(fs, cv2.IMREAD_UNCHANGED), G["colorize_s_png"])


def test_normalize(G):
    assert np.array_equal(iu.normalize(G["normalize_in"]), G["normalize_out"])
    assert np.array_equal(iu.normalize(G["normalize_small_in"]), G["normalize_small_out"])


def test_io_errors_match_reference(golden_dir, tmp_path):
    want = dict()
    for line in open(os.path.join(golden_dir, "reference_errors.txt")).read().splitlines():
        p = line.split("|")
        if p[0] in ("imread", "imwrite"):
            want[p[0]] = p[2]
    with pytest.raises(Exception) as ei:
        iu.imread("/nonexistent/file.png")
    assert str(ei.value) == want["imread"]
    with pytest.raises(Exception) as ei:
        iu.imwrite("/nonexistent_dir/x.png", np.zeros((4, 4, 3), np.uint8))
    assert str(ei.value) == want["imwrite"]


# ---- operator validation happens before any device work ------------------------------------------
def test_apply_filter_validation_matches_reference(golden_dir):
    from reflectance_filtering_b200 import filters
    z = np.zeros((4, 4, 3), np.uint8)
    for line in open(os.path.join(golden_dir, "reference_errors.txt")).read().splitlines():
        p = line.split("|")
        if p[0] in ("imread", "imwrite"):
            continue
        ftype, sc, ss, exc, msg = p
        with pytest.raises(ValueError) as ei:
            filters.apply_filter(ftype, z, z, float(sc), float(ss))
        assert str(ei.value) == msg


def test_blob_shape_check():
    from reflectance_filtering_b200 import cnn
    with pytest.raises(ValueError, match="Expecting to get 1 image in mini-batch having 1 channel"):
        cnn.caffeBlob_to_imgGrayLinear(np.zeros((2, 1, 4, 4), np.float32))
    assert cnn.caffeBlob_to_imgGrayLinear(np.zeros((1, 1, 4, 5), np.float32)).shape == (4, 5)


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from reflectance_filtering_b200 import filters
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        filters.apply_filter("bilateral", np.zeros((8, 8, 3), np.uint8), np.zeros((8, 8, 3), np.uint8), 20, 22)


# ---- CLI surface ------------------------------------------------------------------------------
def test_cli_flags_match_reference():
    import decompose_with_trained_CNN as d
    import filter_reflectance as f
    fopts = {a.dest for a in f.build_parser()._actions}
    assert {"filename_in", "guidance_in", "path_out", "sigma_color", "sigma_spatial", "filter_type"} <= fopts
    a = f.build_parser().parse_args(["--sigma_color=20", "--sigma_spatial=22", "--filter_type=bilateral"])
    assert isinstance(a.sigma_color, float) and a.sigma_spatial == 22.0
    dopts = {a.dest for a in d.build_parser()._actions}
    assert {"filename_in", "path_out"} <= dopts
    assert "_{}_c{}s{}".format("bilateral", a.sigma_color, a.sigma_spatial) == "_bilateral_c20.0s22.0"


def test_cli_help_mode_prints_suggestions():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "filter_reflectance.py")], capture_output=True,
                         text=True, cwd=ROOT, timeout=300).stdout
    assert "--filter_type=bilateral --sigma_color=20 --sigma_spatial=22" in out
    assert "--filter_type=guided --sigma_color=3 --sigma_spatial=45" in out
    out = subprocess.run([sys.executable, os.path.join(ROOT, "decompose_with_trained_CNN.py")],
                         capture_output=True, text=True, cwd=ROOT, timeout=300).stdout
    assert "--filename_in" in out and "--path_out" in out


# ---- C ABI -----------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "rf_b200.h")).read()
    declared = set(re.findall(r"\b(rf_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    L = _native.lib()
    for name in declared:
        assert hasattr(L, name), "librf_b200.so does not export %s" % name
    assert declared == set(_native.SIGNATURES), "ctypes table and header disagree"
    assert L.rf_version() >= 100


def test_geometry_helper_matches_survey():
    L = _native.lib()
    r, t = ctypes.c_int(), ctypes.c_int()
    for ss, er, et in [(22.0, 33, 3409), (28.0, 42, 5525), (0.5, 1, 5), (1.0, 2, 13), (3.0, 4, 49)]:
        assert L.rf_joint_bilateral_geometry(ss, -1, ctypes.byref(r), ctypes.byref(t)) == 0
        assert (r.value, t.value) == (er, et), (ss, r.value, t.value)
    assert L.rf_joint_bilateral_geometry(22.0, 9, ctypes.byref(r), ctypes.byref(t)) == 0 and r.value == 4


def test_guided_row_plan_equals_brute_force_window_sums():
    """Pass B of the guided filter sums a few rows of segment-restarted vertical prefix sums instead of sliding a
    window (csrc/gf2.cu row_terms): check the decomposition against the BORDER_REFLECT window sum it stands for."""
    L = _native.lib()
    rows, wgts = (ctypes.c_int * 12)(), (ctypes.c_float * 12)()
    rng = np.random.default_rng(5)
    worst = 0
    for h, r, seg in [(384, 45, 96), (384, 45, 64), (384, 45, 384), (2160, 45, 64), (100, 64, 100), (65, 64, 64),
                      (97, 45, 65), (50, 8, 50), (200, 33, 67), (46, 45, 46), (130, 64, 65), (9, 8, 9)]:
        v = rng.integers(0, 1000, h).astype(np.int64)
        pv = np.zeros(h, np.int64)                       # prefix sums restarted every `seg` rows, as pass A stores them
        for y in range(h):
            pv[y] = v[y] + (pv[y - 1] if y % seg else 0)
        idx = np.arange(-r, h + r)
        refl = np.where(idx < 0, -idx - 1, np.where(idx >= h, 2 * h - 1 - idx, idx))   # fedcba|abcdefgh|hgfedcb
        assert refl.min() >= 0 and refl.max() < h        # single reflection: h > r
        for y in range(h):
            n = L.rf_guided_row_terms(h, r, seg, y, rows, wgts)
            assert 1 <= n <= 12, (h, r, seg, y, n)
            worst = max(worst, n)
            got = sum(int(wgts[i]) * int(pv[rows[i]]) for i in range(n))
            assert got == int(v[refl[y:y + 2 * r + 1]].sum()), (h, r, seg, y)
    assert worst <= 12
    assert L.rf_guided_row_terms(10, 3, 0, 0, rows, wgts) == -1


def test_argument_errors_do_not_need_a_gpu():
    L = _native.lib()
    assert L.rf_joint_bilateral_u8(None, 3, None, 3, None, 1, 4, 4, 20.0, 22.0, -1, 0, None) == _native.RF_EINVAL
    assert b"NULL" in L.rf_last_error()
    assert L.rf_guided_workspace_bytes(1, 2, 10, 10, 3) >= 2 * 10 * 10 * 16   # coefficient planes (+ padded copies)
    assert L.rf_guided_workspace_bytes(3, 2, 10, 10, 3) >= 3 * L.rf_guided_workspace_bytes(1, 2, 10, 10, 3) // 2
    assert L.rf_guided_workspace_bytes(2, 2, 10, 10, 3) == 0


# ---- sharding ---------------------------------------------------------------------------------
def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 64, 5230):
        for world in (1, 2, 4, 8):
            cover = []
            for r in range(world):
                lo, hi = pipeline.shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                cover += list(range(lo, hi))
            assert cover == list(range(n))
            sizes = [pipeline.shard_range(n, r, world) for r in range(world)]
            assert max(h - l for l, h in sizes) - min(h - l for l, h in sizes) <= 1


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from reflectance_filtering_b200 import pipeline, synth
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
n = 5
lo, hi = pipeline.shard_range(n, rank, world)
# every rank generates only its shard from the per-image seeds; the union must equal the 1-rank batch
mine = synth.batch("stress", hi - lo, 8, 12, 9, start=lo) if hi > lo else np.zeros((0, 8, 12, 3), np.uint8)
chk = torch.tensor([float(mine.astype(np.float64).sum()), float(hi - lo)], dtype=torch.float64)
dist.all_reduce(chk)   # the only collective a run needs: aggregate statistics
full = synth.batch("stress", n, 8, 12, 9)
assert chk[1].item() == n and chk[0].item() == float(full.astype(np.float64).sum()), chk
assert np.array_equal(mine, full[lo:hi])
# mean WHDR over all ranks' images (whdr.reduce_mean): shards of unequal size, weighted by image count
from reflectance_filtering_b200 import whdr
vals = torch.arange(1, n + 1, dtype=torch.float64) / 10.0
mean, count = whdr.reduce_mean(vals[lo:hi])
assert count == n and abs(mean - float(vals.mean())) < 1e-15, (mean, count)
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_sharding_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_whdr_host_helpers():
    """Blob layout and the host-side view of the comparisons (no device needed)."""
    from reflectance_filtering_b200 import whdr
    blob = synth.comparisons(3, 50, seed=5, min_count=10)
    assert blob.shape == (3, 51, 1, 6) and blob.dtype == np.float64
    for b in range(3):
        count = int(blob[b, -1, 0, 0])
        assert 10 <= count <= 50 and np.isnan(blob[b, count:50]).all() and np.isfinite(blob[b, :count]).all()
        comps, name = whdr.get_comparisons_from_blob(blob[b], 30, 40, 0.1)
        assert comps.shape == (count, 6) and name == 100000 + b
        assert (comps[:, [0, 2]] < 40).all() and (comps[:, [1, 3]] < 30).all() and (comps[:, :4] == np.floor(comps[:, :4])).all()
    import torch
    mean, count = whdr.reduce_mean(torch.tensor([0.2, 0.4], dtype=torch.float64))
    assert count == 2 and abs(mean - 0.3) < 1e-15
    assert whdr.reduce_mean(torch.zeros(0, dtype=torch.float64)) == (0.0, 0)


def test_bench_reference_arm_emits_the_contract_line():
    """bench.py --impl reference: exactly one JSON line on stdout with the contract's keys, the same `config` dict the
    CUDA arm builds, and the parity statement (guided filter: unpinned vs ximgproc) that every record carries."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "parity"):
        assert key in rec, key
    assert rec["impl"] == "reference" and rec["gpu_launches"] == 0 and rec["value"] > 0
    assert rec["cpu_baseline"]["kind"] == "port" and rec["cpu_baseline"]["cores"] >= 1
    assert rec["e2e"]["h2d_bytes_per_step"] == 0 and rec["e2e"]["d2h_bytes_per_step"] == 0
    assert "unpinned vs ximgproc" in rec["parity"]["gf"]
    sys.path.insert(0, root)
    import bench
    assert rec["config"] == bench.make_config(64, 1)
