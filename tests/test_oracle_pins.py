"""Pins the CPU oracle (oracle/) to the golden fixtures and to the in-image anchors.

Fixtures in tests/golden/golden.npz come from the reference's own image_utils.py and from
cv2.dnn / cv2.bilateralFilter / cv2.boxFilter (tests/golden/make_golden.py).  Runs without a GPU.
"""
import os

import numpy as np
import pytest

import oracle
from oracle import anchors
from reflectance_filtering_b200 import synth
from reflectance_filtering_b200.caffe_model import DEFAULT_CAFFEMODEL, DEFAULT_PROTOTXT


@pytest.fixture(scope="module")
def G(golden_dir):
    import os
    return np.load(os.path.join(golden_dir, "golden.npz"))


def test_srgb_lut_matches_reference_image_utils(G):
    assert np.array_equal(oracle.srgb_lut(), G["srgb_lut_f32"])
    # SURVEY B.3 spot values
    lut = oracle.srgb_lut()
    assert lut[0] == 0 and lut[255] == 1
    np.testing.assert_allclose(lut[[1, 10, 11, 128]], [3.0352699e-4, 3.0352699e-3, 3.3465358e-3, 0.2158605],
                               rtol=2e-7)


def test_mlp_known_answers(G, mlp):
    # solid colours through the reference input transform, cv2.dnn on the real model
    exp_q = [122, 247, 208, 158, 254, 254, 174, 148]
    for bgr, r_ref, q in zip(G["cnn_solid_bgr"], G["cnn_solid_r"], exp_q):
        r = oracle.mlp_forward(mlp, np.tile(bgr, (4, 4, 1)))
        assert abs(float(r[0, 0]) - float(r_ref)) < 2e-6
        assert int(oracle.quantize_trunc(r)[0, 0]) == q


@pytest.mark.parametrize("key", ["cnn_stress", "cnn_natural"])
def test_mlp_matches_cv2_dnn_golden(G, mlp, key):
    img = G[key + "_in"]
    r = oracle.mlp_forward(mlp, img)
    assert np.abs(r - G[key + "_r"]).max() < 5e-6
    assert np.abs(r - oracle.mlp_forward_f64(mlp, img)).max() < 5e-6


def test_mlp_matches_cv2_dnn_live(mlp):
    img = synth.stress(64, 48, 11)
    live = anchors.DnnNet(DEFAULT_PROTOTXT, DEFAULT_CAFFEMODEL).forward(img)
    assert np.abs(oracle.mlp_forward(mlp, img) - live).max() < 5e-6


def test_weight_reader_matches_cv2_dnn(mlp):
    params = anchors.DnnNet(DEFAULT_PROTOTXT, DEFAULT_CAFFEMODEL).layer_params()
    for i in range(5):
        w, b = params["conv%d" % i]
        assert np.array_equal(w.reshape(mlp.hidden[i][0].shape), mlp.hidden[i][0])
        assert np.array_equal(b.reshape(-1), mlp.hidden[i][1])
    w, b = params["RS_est_before_sigmoid"]  # cv2.dnn names the fusing layer after its top
    assert np.array_equal(w.reshape(-1), mlp.fuse_w)
    assert float(b.reshape(-1)[0]) == mlp.fuse_b


def test_quantize_is_truncation(G):
    assert np.array_equal(oracle.quantize_trunc(G["imwrite_gray_in"]), G["imwrite_gray_png"])
    assert np.array_equal(anchors.quantize_like_imwrite(G["imwrite_gray_in"]), G["imwrite_gray_png"])
    rep = G["imread_gray_png"]
    assert rep.shape[2] == 3 and all(np.array_equal(rep[:, :, c], G["imwrite_gray_png"]) for c in range(3))


@pytest.mark.parametrize("key,sc,ss", [("bf_c20_s22", 20, 22), ("bf_c15_s28", 15, 28), ("bf_c8_s3", 8, 3)])
def test_joint_bilateral_bit_exact_vs_cv2(G, key, sc, ss):
    img = G["bf_in"]
    assert np.array_equal(oracle.joint_bilateral(img.copy(), img, -1, sc, ss), G[key])


def test_joint_bilateral_gray_and_stress(G):
    g = G["bf_gray_in"]
    out = oracle.joint_bilateral(g.copy(), g, -1, 20, 22)
    assert np.array_equal(out, G["bf_gray_c20_s22"])
    assert np.array_equal(out[:, :, 0], out[:, :, 1]) and np.array_equal(out[:, :, 0], out[:, :, 2])
    # a true 1-channel joint is NOT the replicated case: alpha is |dJ|, not 3|dJ|
    one = oracle.joint_bilateral(g[:, :, 0], g[:, :, 0], -1, 20, 22)
    assert not np.array_equal(one, out[:, :, 0])
    # ... it is the replicated case with sigma_color / 3 (same weights up to the rounding of the LUT)
    third = oracle.joint_bilateral(g[:, :, 0], g[:, :, 0], -1, 20 / 3.0, 22)
    assert np.abs(third.astype(int) - out[:, :, 0].astype(int)).max() <= 1
    assert (third != out[:, :, 0]).mean() < 1e-3
    s = G["bf_stress_in"]  # image smaller than the radius: multi-reflection borders
    assert np.array_equal(oracle.joint_bilateral(s.copy(), s, -1, 20, 22), G["bf_stress_c20_s22"])


def test_joint_bilateral_live_anchor_and_properties():
    img = synth.natural(40, 56, 5)
    assert np.array_equal(oracle.joint_bilateral(img.copy(), img, -1, 20, 22),
                          anchors.bilateral_self(img, 20, 22))
    const = np.full((20, 30, 3), 93, np.uint8)
    assert np.array_equal(oracle.joint_bilateral(img[:20, :30], const, -1, 20, 22), const)
    # colour joint + gray-replicated src -> equal output channels
    gsrc = np.repeat(img[:, :, :1], 3, axis=2)
    o = oracle.joint_bilateral(img, gsrc, -1, 20, 22)
    assert np.array_equal(o[:, :, 0], o[:, :, 1]) and np.array_equal(o[:, :, 0], o[:, :, 2])


@pytest.mark.parametrize("jc,sc,scol,ss,d", [(3, 3, 20.0, 6.0, -1), (3, 1, 12.0, 4.0, -1), (1, 3, 30.0, 3.0, 9),
                                              (1, 1, 8.0, 5.0, -1)])
def test_joint_bilateral_distinct_joint_matches_numpy_restatement(jc, sc, scol, ss, d):
    # joint != src is the one bilateral case cv2.bilateralFilter cannot anchor: a second, independently written
    # restatement (numpy on cv2.copyMakeBorder, float32, same tap order) must give the same bytes
    joint = synth.natural(37, 45, 71)
    src = synth.stress(37, 45, 72)
    joint = joint if jc == 3 else np.ascontiguousarray(joint[:, :, 1])
    src = src if sc == 3 else np.ascontiguousarray(src[:, :, 2])
    a = oracle.joint_bilateral(joint, src, d, scol, ss).reshape(src.shape)
    b = anchors.joint_bilateral_numpy(joint, src, d, scol, ss)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("border", [0, 1, 2, 4])
def test_joint_bilateral_border_types_bit_exact_vs_cv2(border):
    """borderType argument: cv2.bilateralFilter takes the same argument, so joint == src pins every type it
    accepts (it rejects BORDER_WRAP, which is checked against the numpy restatement below)."""
    import cv2
    img = synth.natural(41, 37, 81)
    for im in (img, np.ascontiguousarray(img[:, :, 1])):
        want = cv2.bilateralFilter(im, -1, 20.0, 5.0, borderType=border)
        assert np.array_equal(oracle.joint_bilateral(im.copy(), im, -1, 20.0, 5.0, border_type=border), want)


@pytest.mark.parametrize("border", [0, 1, 2, 3, 4])
def test_joint_bilateral_border_types_distinct_joint(border):
    joint, src = synth.natural(23, 31, 82), synth.stress(23, 31, 83)
    a = oracle.joint_bilateral(joint, src, -1, 20.0, 8.0, border_type=border)     # radius 12 > half the height
    assert np.array_equal(a, anchors.joint_bilateral_numpy(joint, src, -1, 20.0, 8.0, border_type=border))


@pytest.mark.parametrize("cn,sc,ss,scale", [(3, 20.0, 4.0, 255.0), (1, 0.1, 3.0, 1.0), (3, 0.05, 2.0, 1.0)])
def test_joint_bilateral_f32_pinned_vs_cv2(cn, sc, ss, scale):
    """CV_32F (SURVEY A.2 last bullet, 8f-4): for joint == src the restatement must agree with cv2.bilateralFilter on
    float32 images (imgproc's bilateralFilter_32f has the structure ximgproc's joint version copies) to float
    rounding: summation order differs (SIMD), nothing else."""
    import cv2
    rng = np.random.default_rng(84)
    img = (rng.random((35, 29, cn)) * scale).astype(np.float32)
    img = img if cn == 3 else img[:, :, 0].copy()
    want = cv2.bilateralFilter(img, -1, sc, ss)
    got = oracle.joint_bilateral(img.copy(), img, -1, sc, ss)
    assert got.dtype == np.float32 and got.shape == want.shape
    assert np.abs(got - want).max() <= 2e-6 * scale


@pytest.mark.parametrize("jc,sc_,border", [(3, 3, 4), (1, 3, 1), (3, 1, 2), (1, 1, 0), (3, 3, 3)])
def test_joint_bilateral_f32_distinct_joint_matches_numpy_restatement(jc, sc_, border):
    rng = np.random.default_rng(85)
    joint = (rng.random((27, 33, 3)) * 255).astype(np.float32)
    src = (rng.random((27, 33, 3)) * 100 - 20).astype(np.float32)
    joint = joint if jc == 3 else joint[:, :, 0].copy()
    src = src if sc_ == 3 else src[:, :, 1].copy()
    a = oracle.joint_bilateral(joint, src, -1, 25.0, 3.0, border_type=border)
    b = anchors.joint_bilateral_f32_numpy(joint, src, -1, 25.0, 3.0, border_type=border)
    assert a.shape == src.shape and np.abs(a - b).max() <= 1e-4 * 120      # float32 sums in the same order: ~1 ulp
    flat = np.full_like(joint, 7.5)                                         # constant joint: spatial Gaussian only
    c = oracle.joint_bilateral(flat, src, -1, 25.0, 3.0, border_type=border)
    assert np.isfinite(c).all()


@pytest.mark.parametrize("r", [1, 7, 45])
def test_box_mean_bit_exact_vs_cv2(G, r):
    assert np.array_equal(oracle.box_mean_reflect(G["box_in"], r), G["box_r%d" % r])


@pytest.mark.parametrize("r,eps,sc", [(45, 3.0, 3), (7, 3.0, 3), (52, 7.0, 1), (2, 0.5, 3)])
def test_guided_matches_cv2box_restatement(r, eps, sc):
    # the C restatement and the numpy-on-cv2.boxFilter restatement are independent codes of
    # SURVEY A.3; parity against a real ximgproc build remains unpinned (not installable)
    gd = synth.flat(72, 60, 31)
    src = synth.natural(72, 60, 32)
    if sc == 1:
        src = src[:, :, 0]
    a = oracle.guided(gd, src, r, eps)
    b = anchors.guided_cv2box(gd, src, r, eps)
    assert np.abs(a.astype(int) - b.astype(int)).max() <= 1
    assert (a != b).mean() < 1e-3


@pytest.mark.parametrize("r,eps,sc", [(9, 3.0, 1), (45, 7.0, 3), (2, 0.5, 1)])
def test_guided_gray_guide_matches_cv2box_restatement(r, eps, sc):
    # 1-channel guide (the rest of the cv2.ximgproc.guidedFilter surface; not reachable from the reference CLI):
    # two independent restatements, and the textbook identity GF_gray(I, p, eps) == p when p == I and eps -> 0
    gd = synth.flat(64, 70, 41)[:, :, 1].copy()
    src = synth.natural(64, 70, 42)
    if sc == 1:
        src = src[:, :, 0].copy()
    a = oracle.guided(gd, src, r, eps)
    b = anchors.guided_cv2box(gd, src, r, eps)
    assert a.shape == src.shape and np.abs(a.astype(int) - b.astype(int)).max() <= 1
    assert (a != b).mean() < 1e-3
    self_guided = oracle.guided(gd, gd, r, 1e-4)
    assert np.abs(self_guided.astype(int) - gd.astype(int)).max() <= 1


@pytest.mark.parametrize("h,w,r,eps,sc,gc", [(96, 120, 45, 3.0, 3, 3), (72, 60, 7, 3.0, 3, 3), (64, 64, 52, 7.0, 1, 3),
                                             (30, 200, 45, 3.0, 1, 3), (5, 3, 2, 0.5, 3, 3), (1, 1, 1, 1.0, 3, 3),
                                             (96, 120, 45, 3.0, 1, 1), (50, 90, 20, 0.5, 1, 1), (120, 160, 45, 3.0, 1, 3)])
def test_guided_bounded_by_float64_paper_formulation(h, w, r, eps, sc, gc):
    """Third formulation (float64, np.linalg.solve, summed-area table): the restated guided_filter.cpp arithmetic must
    stay within 1 LSB of it (measured: < 1e-4 of the bytes) on every shape the GPU tests use -- a bound, not a pin (ximgproc itself is not available)."""
    gd = synth.flat(h, w, 51)
    src = synth.natural(h, w, 52)
    if sc == 1:
        src = np.ascontiguousarray(src[:, :, 0])
    if gc == 1:
        gd = np.ascontiguousarray(gd[:, :, 2])
    a = oracle.guided(gd, src, r, eps)
    b = anchors.guided_float64(gd, src, r, eps)
    d = np.abs(a.astype(int) - b.astype(int))
    assert a.shape == b.shape and d.max() <= 1, d.max()       # measured: at most 1 LSB ...
    assert (d > 0).mean() < 1e-3, (d > 0).mean()                   # ... on < 1e-4 of the bytes (rounding ties)


@pytest.mark.parametrize("gdt,sdt,sc,gc", [("f", "f", 3, 3), ("f", "f", 1, 1), ("u", "f", 1, 3), ("f", "u", 3, 3)])
def test_guided_float_depths(gdt, sdt, sc, gc):
    """CV_32F guide and / or source: guidedFilter converts to float without scaling; dst has src's depth."""
    rng = np.random.default_rng(86)
    gd = (rng.random((44, 52, 3)) * 255).astype(np.float32 if gdt == "f" else np.uint8)
    src = (rng.random((44, 52, 3)) * 200).astype(np.float32 if sdt == "f" else np.uint8)
    gd = gd if gc == 3 else np.ascontiguousarray(gd[:, :, 0])
    src = src if sc == 3 else np.ascontiguousarray(src[:, :, 1])
    a = oracle.guided(gd, src, 9, 3.0)
    b = anchors.guided_cv2box(gd, src, 9, 3.0)
    c = anchors.guided_float64(gd, src, 9, 3.0)
    assert a.dtype == src.dtype and a.shape == src.shape
    if sdt == "f":
        assert np.abs(a - b).max() <= 1e-5 and np.abs(a.astype(np.float64) - c).max() <= 1e-3
    else:
        assert np.abs(a.astype(int) - b.astype(int)).max() <= 1 and np.abs(a.astype(int) - c.astype(int)).max() <= 1


def test_guided_properties():
    gd = synth.flat(48, 40, 41)
    const = np.full((48, 40, 3), 77, np.uint8)
    assert np.array_equal(oracle.guided(gd, const, 9, 3.0), const)
    # a gray-replicated source filters to equal channels, each equal to the 1-channel result
    s = synth.natural(48, 40, 42)[:, :, :1]
    o3 = oracle.guided(gd, np.repeat(s, 3, axis=2), 9, 3.0)
    o1 = oracle.guided(gd, s[:, :, 0], 9, 3.0)
    assert all(np.array_equal(o3[:, :, c], o1) for c in range(3))


def test_apply_filter_errors_match_reference(golden_dir):
    import os
    z = np.zeros((4, 4, 3), np.uint8)
    for line in open(os.path.join(golden_dir, "reference_errors.txt")).read().splitlines():
        parts = line.split("|")
        if parts[0] in ("imread", "imwrite"):
            continue
        ftype, sc, ss, exc, msg = parts
        with pytest.raises(ValueError) as ei:
            oracle.apply_filter(ftype, z, z, float(sc), float(ss))
        assert exc == "ValueError" and str(ei.value) == msg


def test_whdr_restatement_matches_reference_code(golden_dir):
    """golden_whdr.npz was produced by the reference's own whdr() (tests/golden/make_golden_whdr.py)."""
    W = np.load(os.path.join(golden_dir, "golden_whdr.npz"))
    for case in W["cases"]:
        refl, blob, delta = W[case + "_reflectance"], W[case + "_blob"], float(W[case + "_delta"])
        got = np.array([oracle.whdr(refl[b], blob[b], delta) for b in range(refl.shape[0])])
        np.testing.assert_array_equal(got, W[case + "_whdr"])
        assert got[0] == 0.0 and (got[1:] > 0).all()
