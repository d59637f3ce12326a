#!/usr/bin/env python
"""Generates the golden fixtures under tests/golden/ (run in the dev container only).

Two kinds of vectors:
  * outputs of the REFERENCE ITSELF: /root/reference/image_utils.py is imported unmodified and
    run (sRGB helpers, imwrite quantisation through a real PNG round trip, colorize, normalize);
    /root/reference/filter_reflectance.py is imported for its error behaviour;
  * outputs of independent third-party code on the reference's real artefacts / algorithm, for
    the parts of the path that live in un-vendored Caffe / OpenCV-contrib: cv2.dnn reading
    network_definition.prototxt + learned_weights.caffemodel, cv2.bilateralFilter, cv2.boxFilter.

/root/reference does not exist on the GPU box; the tests only read the .npz files written here.
"""
import os
import sys
import tempfile

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from reflectance_filtering_b200 import synth  # noqa: E402  (deterministic input generators only)


def load_reference_module(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    riu = load_reference_module("image_utils")
    # numpy 2 renamed percentile's keyword; the reference passes interpolation='lower'
    _pct = np.percentile
    def _percentile(a, q, interpolation=None, **kw):
        if interpolation is not None:
            kw["method"] = interpolation
        return _pct(a, q, **kw)
    riu.np.percentile = _percentile

    out = {}
    # ---- image_utils conventions, from the reference's own code -------------------------------
    x = np.concatenate([np.linspace(0, 1, 513), np.arange(256) / 255.0, [0.04045, 0.0031308, 0.04, 0.05]])
    out["srgb_in"] = x
    out["srgb_to_rgb"] = riu.srgb_to_rgb(x)
    out["rgb_to_srgb"] = riu.rgb_to_srgb(x)
    out["srgb_to_rgb_f32"] = riu.srgb_to_rgb(x.astype(np.float32))
    codes = np.arange(256, dtype=np.float64) / 255.0
    out["srgb_lut_f32"] = riu.srgb_to_rgb(codes).astype(np.float32)

    rng = np.random.default_rng(12345)
    gray = rng.uniform(0.0, 0.999, (24, 20)).astype(np.float32)
    img = synth.natural(24, 20, 77)
    with tempfile.TemporaryDirectory() as td:
        f = os.path.join(td, "g-r.png")
        riu.imwrite(f, gray)
        raw = cv2.imread(f, cv2.IMREAD_UNCHANGED)
        out["imwrite_gray_in"] = gray
        out["imwrite_gray_png"] = raw                 # 1 channel, truncation
        out["imread_gray_png"] = riu.imread(f)        # 3 equal channels
        refl, shad = riu.colorize(gray, img)
        out["colorize_image"] = img
        out["colorize_reflectance"] = refl
        out["colorize_shading"] = shad
        fr, fs = os.path.join(td, "r.png"), os.path.join(td, "s.png")
        riu.imwrite(fr, refl, sRGB=True)
        riu.imwrite(fs, shad, sRGB=True)
        out["colorize_r_png"] = cv2.imread(fr, cv2.IMREAD_UNCHANGED)
        out["colorize_s_png"] = cv2.imread(fs, cv2.IMREAD_UNCHANGED)
    big = rng.uniform(0, 300, (16, 16, 3))
    out["normalize_in"] = big
    out["normalize_out"] = riu.normalize(big)
    small = rng.uniform(0, 0.9, (8, 8))
    out["normalize_small_in"] = small
    out["normalize_small_out"] = riu.normalize(small)

    # ---- CNN: cv2.dnn on the reference's real prototxt + caffemodel ----------------------------
    net = cv2.dnn.readNetFromCaffe(os.path.join(REF, "network_definition.prototxt"),
                                   os.path.join(REF, "learned_weights.caffemodel"))

    def dnn_forward(bgr):
        blob = bgr / 255.0
        blob = blob[:, :, ::-1]
        blob = riu.srgb_to_rgb(blob)
        blob = np.transpose(blob, (2, 0, 1))[np.newaxis].astype(np.float32)
        net.setInput(np.ascontiguousarray(blob))
        return net.forward()[0, 0].copy()

    solid = np.array([(0, 0, 0), (255, 255, 255), (128, 128, 128), (0, 0, 255), (0, 255, 0), (255, 0, 0),
                      (10, 128, 240), (200, 50, 25)], np.uint8)
    out["cnn_solid_bgr"] = solid
    out["cnn_solid_r"] = np.array([dnn_forward(np.tile(c, (4, 4, 1)))[0, 0] for c in solid], np.float32)
    cin = synth.stress(24, 20, 2001)
    out["cnn_stress_in"] = cin
    out["cnn_stress_r"] = dnn_forward(cin)
    cnat = synth.natural(32, 48, 2002)
    out["cnn_natural_in"] = cnat
    out["cnn_natural_r"] = dnn_forward(cnat)

    # ---- joint bilateral: cv2.bilateralFilter == jointBilateralFilter(joint = src copy) --------
    bimg = synth.natural(48, 40, 1001)
    out["bf_in"] = bimg
    out["bf_c20_s22"] = cv2.bilateralFilter(bimg, -1, 20.0, 22.0)
    out["bf_c15_s28"] = cv2.bilateralFilter(bimg, -1, 15.0, 28.0)
    out["bf_c8_s3"] = cv2.bilateralFilter(bimg, -1, 8.0, 3.0)
    bgray = np.repeat(bimg[:, :, 1:2], 3, axis=2)
    out["bf_gray_in"] = bgray
    out["bf_gray_c20_s22"] = cv2.bilateralFilter(bgray, -1, 20.0, 22.0)
    bst = synth.stress(20, 24, 1002)
    out["bf_stress_in"] = bst
    out["bf_stress_c20_s22"] = cv2.bilateralFilter(bst, -1, 20.0, 22.0)

    # ---- box mean primitive of the guided filter ---------------------------------------------
    pl = (synth.stress(40, 36, 3001)[:, :, 0].astype(np.float32) *
          synth.stress(40, 36, 3002)[:, :, 0].astype(np.float32))
    out["box_in"] = pl
    for r in (1, 7, 45):
        k = 2 * r + 1
        out["box_r%d" % r] = cv2.boxFilter(pl, cv2.CV_32F, (k, k), normalize=True,
                                           borderType=cv2.BORDER_REFLECT)

    np.savez_compressed(os.path.join(HERE, "golden.npz"), **out)
    print("wrote", os.path.join(HERE, "golden.npz"), "with", len(out), "arrays")

    # ---- error behaviour of the reference operator (text fixture) -----------------------------
    rfr = None
    try:
        sys.path.insert(0, REF)
        rfr = load_reference_module("filter_reflectance")
    except Exception as e:  # pragma: no cover
        print("reference filter_reflectance not importable:", e)
    lines = []
    if rfr is not None:
        z = np.zeros((4, 4, 3), np.uint8)
        for args in [("bilateral", z, z, 0, 1), ("bilateral", z, z, 1, -1), ("nope", z, z, 1, 1)]:
            try:
                rfr.apply_filter(*args)
                lines.append("%s|%s|%s|no error" % (args[0], args[3], args[4]))
            except Exception as e:
                lines.append("%s|%s|%s|%s|%s" % (args[0], args[3], args[4], type(e).__name__, e))
        try:
            riu.imread("/nonexistent/file.png")
        except Exception as e:
            lines.append("imread|%s|%s" % (type(e).__name__, e))
        try:
            riu.imwrite("/nonexistent_dir/x.png", z)
        except Exception as e:
            lines.append("imwrite|%s|%s" % (type(e).__name__, e))
    with open(os.path.join(HERE, "reference_errors.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
