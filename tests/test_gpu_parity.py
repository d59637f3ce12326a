"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs, against the committed golden fixtures, and -- at BASELINE.json sizes -- through
size-independent properties.  Tolerances (BASELINE.json north_star): max abs error <= 1e-3 on the
linear reflectance; uint8 outputs within +-1 LSB (we additionally bound the fraction of bytes
that differ at all)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import oracle  # noqa: E402
from reflectance_filtering_b200 import cnn, filters, image_utils as iu, pipeline, synth  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CNN_TOL = 1e-3          # north_star tolerance on linear reflectance
CNN_TIGHT = 2e-5        # what the 3xTF32 tensor-core kernel achieves against the FP32 oracle (FP32 kernel: 2e-6)


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "golden.npz"))


@pytest.fixture(scope="module")
def net():
    return cnn.default_net()


def dev_u8(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def lsb_stats(a, b):
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    return int(d.max()), float((d != 0).mean())


# ---- CNN ---------------------------------------------------------------------------------------
def test_cnn_known_answers_and_golden(G, net):
    for bgr, r_ref in zip(G["cnn_solid_bgr"], G["cnn_solid_r"]):
        r = cnn.get_reflectance_caffe(net, np.tile(bgr, (4, 4, 1)))
        assert r.dtype == np.float32 and r.shape == (4, 4)
        assert np.abs(r - r_ref).max() < CNN_TIGHT
    for key in ("cnn_stress", "cnn_natural"):
        r = cnn.get_reflectance_caffe(net, G[key + "_in"])
        assert np.abs(r - G[key + "_r"]).max() < CNN_TIGHT  # cv2.dnn on the real artefacts


@pytest.mark.parametrize("h,w,kind", [(384, 512, "natural"), (37, 53, "stress"), (1, 1, "stress"), (3, 1, "stress")])
def test_cnn_matches_oracle(net, mlp, h, w, kind):
    img = synth.GENERATORS[kind](h, w, 4242)
    r = cnn.get_reflectance_caffe(net, img)
    ref = oracle.mlp_forward(mlp, img)
    err = np.abs(r - ref).max()
    assert err < CNN_TOL and err < CNN_TIGHT
    assert np.abs(r.astype(np.float64) - oracle.mlp_forward_f64(mlp, img)).max() < CNN_TIGHT


def test_cnn_tensor_core_and_fp32_kernels_agree():
    # the same library, RF_CNN_FP32=1 forces the exact-FP32 CUDA-core kernel; both must sit within the
    # north-star tolerance of the oracle and within 2e-5 of each other
    code = ("import sys, numpy as np; sys.path.insert(0, %r); "
            "from reflectance_filtering_b200 import cnn, synth; "
            "r = cnn.get_reflectance_caffe(cnn.default_net(), synth.stress(200, 333, 99)); "
            "np.save(sys.argv[1], r)" % ROOT)
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        outs = []
        for i, env in enumerate(({}, {"RF_CNN_FP32": "1"})):
            f = os.path.join(td, "r%d.npy" % i)
            subprocess.run([sys.executable, "-c", code, f], check=True, env=dict(os.environ, **env), timeout=600)
            outs.append(np.load(f))
    assert np.abs(outs[0] - outs[1]).max() < CNN_TIGHT
    ref = oracle.mlp_forward(cnn.default_net().mlp, synth.stress(200, 333, 99))
    assert np.abs(outs[1] - ref).max() < 5e-6 and np.abs(outs[0] - ref).max() < CNN_TIGHT


class _SyntheticMLP(object):
    """Random per-pixel MLP in the parameter layout of include/rf_b200.h (what caffe_model.PixelMLP provides)."""

    def __init__(self, width, n_hidden, seed):
        rng = np.random.default_rng(seed)
        self.hidden, k = [], 3
        for _ in range(n_hidden):
            self.hidden.append((rng.normal(0, 1.2 / np.sqrt(k), (width, k)).astype(np.float32),
                                rng.normal(0, 0.3, width).astype(np.float32)))
            k = width
        self.fuse_w = rng.normal(0, 0.5 / np.sqrt(width * n_hidden), width * n_hidden).astype(np.float32)
        self.fuse_b = np.float32(0.1)

    def flat_params(self):
        parts = []
        for w, b in self.hidden:
            parts += [w.reshape(-1), b]
        return np.ascontiguousarray(np.concatenate(parts + [self.fuse_w, np.asarray([self.fuse_b], np.float32)]))

    def dims(self):
        return [3] + [w.shape[0] for w, _ in self.hidden]


@pytest.mark.parametrize("width,n_hidden", [(32, 2), (32, 3), (32, 8), (16, 4)])
def test_cnn_other_depths_and_widths_through_the_c_abi(width, n_hidden):
    """rf_cnn_create / rf_cnn_forward_u8 with graphs other than the shipped one: width 32 runs on the tensor
    cores for 2..8 hidden layers (the bias row and conv0 planes are built per layer count), other widths on the
    exact-FP32 kernel; both against the CPU oracle on a stress image with an odd pixel count."""
    import ctypes as C
    from reflectance_filtering_b200 import _native, device as dev
    m = _SyntheticMLP(width, n_hidden, 17 * width + n_hidden)
    params, dims = m.flat_params(), np.asarray(m.dims(), np.int32)
    lut = np.ascontiguousarray(iu.srgb_lut(), np.float32)
    handle = C.c_void_p()
    _native.check(_native.lib().rf_cnn_create(params.ctypes.data_as(C.c_void_p), dims.ctypes.data_as(C.c_void_p),
                                              len(dims) - 1, lut.ctypes.data_as(C.c_void_p), C.byref(handle)))
    try:
        img = synth.stress(61, 83, 5)
        d = dev_u8(img[None])
        out = torch.empty((1, 61, 83), dtype=torch.float32, device="cuda")
        _native.check(_native.lib().rf_cnn_forward_u8(handle, dev.ptr(d), 1, 61, 83, dev.ptr(out), C.c_void_p(0),
                                                      dev.stream_ptr()))
        got = out.cpu().numpy()[0]
    finally:
        _native.lib().rf_cnn_destroy(handle)
    ref = oracle.mlp_forward(m, img)
    err = float(np.abs(got - ref).max())
    assert err < CNN_TOL and err < 1e-4, err   # random weights: larger activations than the shipped model


def test_cnn_fused_quantisation_is_truncation(net, mlp):
    imgs = synth.batch("natural", 3, 96, 80, 2)
    f32, u8 = net.forward_device(dev_u8(imgs), want_f32=True, want_u8=True)
    f32, u8 = f32.cpu().numpy(), u8.cpu().numpy()
    assert np.array_equal(u8, (f32 * 255).astype(np.uint8))          # image_utils.py:68 on our floats
    ref = np.stack([oracle.quantize_trunc(oracle.mlp_forward(mlp, im)) for im in imgs])
    mx, frac = lsb_stats(u8, ref)
    assert mx <= 1 and frac < 1e-3                                    # only boundary straddlers differ
    # batch == per-image, odd pixel counts handled
    odd = synth.batch("stress", 2, 5, 7, 3)
    a = net.forward_device(dev_u8(odd))[0].cpu().numpy()
    b = np.stack([cnn.get_reflectance_caffe(net, im) for im in odd])
    assert np.array_equal(a, b)


# ---- joint bilateral ---------------------------------------------------------------------------
@pytest.mark.parametrize("key,sc,ss", [("bf_c20_s22", 20, 22), ("bf_c15_s28", 15, 28), ("bf_c8_s3", 8, 3)])
def test_bf_golden_self_guided(G, key, sc, ss):
    img = G["bf_in"]
    out = filters.apply_filter("bilateral", img, img.copy(), sc, ss)
    mx, frac = lsb_stats(out, G[key])                                # cv2.bilateralFilter bytes
    assert out.shape == img.shape and out.dtype == np.uint8
    assert mx <= 1 and frac < 2e-3, (mx, frac)
    same_obj = filters.apply_filter("bilateral", img, img, sc, ss)   # aliased joint: one staged window
    assert np.array_equal(same_obj, out)


def test_bf_golden_gray_and_small(G):
    g = G["bf_gray_in"]
    out = filters.apply_filter("bilateral", g, g.copy(), 20, 22)     # takes the single-channel kernel
    mx, frac = lsb_stats(out, G["bf_gray_c20_s22"])
    assert mx <= 1 and frac < 2e-3, (mx, frac)
    assert np.array_equal(out[:, :, 0], out[:, :, 1]) and np.array_equal(out[:, :, 0], out[:, :, 2])
    s = G["bf_stress_in"]                                            # 20x24 image, radius 33
    mx, frac = lsb_stats(filters.apply_filter("bilateral", s, s.copy(), 20, 22), G["bf_stress_c20_s22"])
    assert mx <= 1 and frac < 5e-3, (mx, frac)


@pytest.mark.parametrize("h,w", [(96, 80), (33, 131), (8, 300), (1, 1), (2, 5)])
def test_bf_joint_differs_from_src(h, w):
    joint = synth.natural(h, w, 10)
    src = synth.stress(h, w, 11)
    out = filters.apply_filter("bilateral", src, joint, 20, 22)
    mx, frac = lsb_stats(out, oracle.joint_bilateral(joint, src, -1, 20, 22))
    assert mx <= 1 and frac < 2e-3, (mx, frac)


def test_bf_channel_combinations():
    joint = synth.natural(40, 56, 20)
    src = synth.natural(40, 56, 21)
    dj, ds = dev_u8(joint[None]), dev_u8(src[None])
    for jt, st, oj, os_ in [(dj[..., :1].contiguous(), ds, joint[:, :, :1], src),
                            (dj, ds[..., 0].contiguous(), joint, src[:, :, 0]),
                            (dj[..., 0].contiguous(), ds[..., 0].contiguous(), joint[:, :, 0], src[:, :, 0])]:
        out = filters.joint_bilateral_device(jt, st, 20, 22).cpu().numpy()[0]
        ref = oracle.joint_bilateral(oj, os_, -1, 20, 22)
        mx, frac = lsb_stats(out.reshape(ref.shape), ref)
        assert mx <= 1 and frac < 2e-3, (mx, frac)
    # explicit d overrides the radius
    out = filters.joint_bilateral_device(dj, ds, 20, 22, d=9).cpu().numpy()[0]
    mx, frac = lsb_stats(out, oracle.joint_bilateral(joint, src, 9, 20, 22))
    assert mx <= 1 and frac < 2e-3


def test_bf_full_size_config1_properties_and_oracle():
    # BASELINE config 1: 512x384, c20 s22, a copy of itself as joint
    img = synth.natural(384, 512, 1000)
    out = filters.apply_filter("bilateral", img, img.copy(), 20, 22)
    ref = oracle.joint_bilateral(img.copy(), img, -1, 20, 22)
    mx, frac = lsb_stats(out, ref)
    assert mx <= 1 and frac < 1e-3, (mx, frac)
    const = np.full((384, 512, 3), 200, np.uint8)                    # constants are fixed points
    assert np.array_equal(filters.apply_filter("bilateral", const, img, 20, 22), const)
    gsrc = np.repeat(img[:, :, :1], 3, axis=2)                       # equal src channels stay equal
    o = filters.apply_filter("bilateral", gsrc, img, 20, 22)
    assert np.array_equal(o[:, :, 0], o[:, :, 1]) and np.array_equal(o[:, :, 0], o[:, :, 2])


def test_bf_batch_equals_per_image_and_radius_limit():
    imgs = synth.batch("natural", 3, 64, 72, 6)
    d = dev_u8(imgs)
    batch = filters.joint_bilateral_device(d, d, 15, 28).cpu().numpy()
    for i in range(3):
        assert np.array_equal(batch[i], filters.apply_filter("bilateral", imgs[i], imgs[i], 15, 28))


@pytest.mark.parametrize("jc,sc,r,gray_rep", [(3, 3, 70, False), (1, 1, 100, True), (3, 1, 64, False), (1, 3, 66, False)])
def test_bf_large_radius_runs_on_the_generic_kernel(jc, sc, r, gray_rep):
    """Radii beyond the shared-memory-tiled kernels (rf_joint_bilateral_fast_max_radius(), and r = 61..64 with a
    distinct colour joint) no longer fail: the generic kernel takes them, bit-equal to the oracle."""
    from reflectance_filtering_b200 import _native
    assert _native.lib().rf_joint_bilateral_fast_max_radius() == 64 and _native.lib().rf_joint_bilateral_max_radius() >= 1024
    joint, src = synth.natural(50, 60, 91), synth.stress(50, 60, 92)
    joint = joint if jc == 3 else np.ascontiguousarray(joint[:, :, 1])
    src = src if sc == 3 else np.ascontiguousarray(src[:, :, 2])
    if gray_rep:
        src = joint
    out = filters.joint_bilateral_device(dev_u8(joint[None]), dev_u8(src[None]), 20.0, 10.0, d=2 * r + 1,
                                         gray_replicated=gray_rep).cpu().numpy()[0]
    if gray_rep:
        j3 = np.repeat(joint[:, :, None], 3, axis=2)
        ref = oracle.joint_bilateral(j3, j3, 2 * r + 1, 20.0, 10.0)[:, :, 0]
    else:
        ref = oracle.joint_bilateral(joint, src, 2 * r + 1, 20.0, 10.0)
    assert np.array_equal(out, ref.reshape(out.shape))
    with pytest.raises(_native.NativeError, match="radius"):
        filters.joint_bilateral_device(dev_u8(joint[None]), dev_u8(src[None]), 20, 2000.0)


@pytest.mark.parametrize("border", [0, 1, 2, 3, 4])
def test_bf_border_types(border):
    joint, src = synth.natural(33, 47, 93), synth.stress(33, 47, 94)
    for (j, s_) in ((joint, src), (np.ascontiguousarray(joint[:, :, 0]), np.ascontiguousarray(src[:, :, 0]))):
        out = filters.joint_bilateral_device(dev_u8(j[None]), dev_u8(s_[None]), 20.0, 6.0, border_type=border).cpu().numpy()[0]
        ref = oracle.joint_bilateral(j, s_, -1, 20.0, 6.0, border_type=border)
        if border == 4:    # the tiled kernels: +-1 LSB of rounding ties
            mx, frac = lsb_stats(out, ref.reshape(out.shape))
            assert mx <= 1 and frac < 2e-3
        else:              # the generic kernel repeats the oracle's arithmetic
            assert np.array_equal(out, ref.reshape(out.shape)), border


@pytest.mark.parametrize("jc,sc,border", [(3, 3, 4), (1, 1, 4), (3, 1, 1), (1, 3, 0), (3, 3, 3)])
def test_bf_float32_images(jc, sc, border):
    """CV_32F joint / src (jointBilateralFilter_32f) through rf_joint_bilateral_f32 and the numpy operator."""
    rng = np.random.default_rng(95)
    n = 2
    joint = (rng.random((n, 40, 52, 3)) * 255).astype(np.float32)
    src = (rng.random((n, 40, 52, 3)) * 300 - 50).astype(np.float32)
    joint = joint if jc == 3 else np.ascontiguousarray(joint[..., 0])
    src = src if sc == 3 else np.ascontiguousarray(src[..., 1])
    out = filters.joint_bilateral_device(torch.from_numpy(joint).cuda(), torch.from_numpy(src).cuda(), 30.0, 4.0,
                                         border_type=border).cpu().numpy()
    for i in range(n):   # the exp table is scaled to each image's own range
        ref = oracle.joint_bilateral(joint[i], src[i], -1, 30.0, 4.0, border_type=border)
        assert np.abs(out[i] - ref.reshape(out[i].shape)).max() <= 1e-3, i     # values up to 250: a few ulp
    if border == 4:
        got = filters.apply_filter("bilateral", src[0], joint[0], 30.0, 4.0)
        assert got.dtype == np.float32 and np.abs(got - out[0].reshape(got.shape)).max() == 0
        self_guided = filters.apply_filter("bilateral", joint[0], joint[0], 30.0, 4.0)
        import cv2
        want = cv2.bilateralFilter(joint[0], -1, 30.0, 4.0)
        assert np.abs(self_guided - want).max() <= 2e-6 * 255 * 2
    with pytest.raises(TypeError):
        filters.apply_filter("bilateral", src[0], joint[0].astype(np.uint8), 30.0, 4.0)


# ---- guided filter -----------------------------------------------------------------------------
@pytest.mark.parametrize("h,w,r,eps,sc", [(96, 120, 45, 3.0, 3), (72, 60, 7, 3.0, 3), (64, 64, 52, 7.0, 1),
                                          (30, 700, 45, 3.0, 1), (5, 3, 2, 0.5, 3), (1, 1, 1, 1.0, 3)])
def test_gf_matches_oracle(h, w, r, eps, sc):
    gd = synth.flat(h, w, 51)
    src = synth.natural(h, w, 52)
    if sc == 1:
        src = src[:, :, 0]
    out = filters.apply_filter("guided", src, gd, eps, float(r) + 0.7)   # radius = int(sigma_spatial)
    ref = oracle.guided(gd, src, r, eps)
    mx, frac = lsb_stats(out, ref)
    assert out.shape == ref.shape and mx <= 1 and frac < 2e-3, (mx, frac)


@pytest.mark.parametrize("h,w,r,eps,sc", [(96, 120, 45, 3.0, 1), (72, 60, 7, 3.0, 3), (50, 610, 20, 0.5, 1), (1, 1, 1, 1.0, 3)])
def test_gf_gray_guide_matches_oracle(h, w, r, eps, sc):
    # 1-channel guide: device entry point and numpy operator (2-D joint), against the oracle; iterated == repeated
    gd = np.ascontiguousarray(synth.flat(h, w, 61)[:, :, 2])
    src = synth.natural(h, w, 62)
    src = src if sc == 3 else np.ascontiguousarray(src[:, :, 0])
    ref = oracle.guided(gd, src, r, eps)
    out = filters.apply_filter("guided", src, gd, eps, float(r) + 0.3)
    assert out.shape == ref.shape and np.array_equal(out, ref)      # the generic kernels are bit-equal to the oracle
    dg, ds = dev_u8(gd[None]), dev_u8(src[None])
    twice = filters.guided_device(dg, filters.guided_device(dg, ds, r, eps), r, eps)
    assert torch.equal(filters.guided_device(dg, ds, r, eps, iterations=2), twice)


@pytest.mark.parametrize("h,w,r,sc", [(33, 47, 7, 3), (33, 47, 7, 1), (40, 57, 12, 1), (64, 100, 20, 3)])
def test_gf_ignores_workspace_contents(h, w, r, sc):
    # the workspace is scratch: results must not depend on what it held (regression: inf - inf in a prefix)
    gd = synth.flat(h, w, 802)
    src = synth.natural(h, w, 702)
    src = src if sc == 3 else src[:, :, 0].copy()
    ref = oracle.guided(gd, src, r, 3.0)
    for fill in (float("nan"), 3e38, -3e38):
        ws = torch.full((8 << 20,), fill, dtype=torch.float32, device="cuda").view(torch.uint8)
        out = filters.guided_device(dev_u8(gd[None]), dev_u8(src[None]), r, 3.0, workspace=ws).cpu().numpy()[0]
        mx, frac = lsb_stats(out, ref)
        assert mx <= 1 and frac < 2e-3, (fill, mx, frac)


@pytest.mark.parametrize("r", [127, 128, 150, 240])
def test_gf_large_radius_bright_image(r):
    """Window sums of I*I' on a bright image exceed 2^32 from r = 128 on (65,025 * 257^2): the generic kernels switch
    to 64-bit sums there (uint32 sums silently wrapped).  Bit-equal to the oracle on both sides."""
    h, w = 2 * r + 40, 300
    rng = np.random.default_rng(r)
    gd = (255 - rng.integers(0, 6, size=(h, w, 3))).astype(np.uint8)       # saturated guide with a little texture
    src = (250 - rng.integers(0, 40, size=(h, w))).astype(np.uint8)
    out = filters.guided_device(dev_u8(gd[None]), dev_u8(src[None]), r, 3.0).cpu().numpy()[0]
    ref = oracle.guided(gd, src, r, 3.0)
    mx, frac = lsb_stats(out, ref)
    assert mx <= 1 and frac < 2e-3, (r, mx, frac)
    from reflectance_filtering_b200 import _native
    assert _native.lib().rf_guided_max_radius() >= 240


def test_gf_bytes_do_not_depend_on_batch_composition():
    """An image's output must be the same bytes whether it is filtered alone, in a chunk or in the whole batch
    (sharded == unsharded).  The row segmentation of the fast path once depended on the batch size."""
    n, h, w, r = 12, 200, 300, 20
    gd = np.stack([synth.flat(h, w, 900 + i) for i in range(n)])
    src = np.ascontiguousarray(np.stack([synth.natural(h, w, 950 + i) for i in range(n)])[..., 1])
    dg, ds = dev_u8(gd), dev_u8(src)
    for iters in (1, 3):
        whole = filters.guided_device(dg, ds, r, 3.0, iterations=iters)
        for lo, hi in ((0, 1), (1, 4), (4, 12)):
            part = filters.guided_device(dg[lo:hi].contiguous(), ds[lo:hi].contiguous(), r, 3.0, iterations=iters)
            assert torch.equal(part, whole[lo:hi]), (iters, lo, hi)


@pytest.mark.parametrize("gdt,sdt,sc,gc,r", [("f", "f", 3, 3, 9), ("f", "f", 1, 1, 20), ("u", "f", 1, 3, 45), ("f", "u", 3, 3, 7)])
def test_gf_float_depths(gdt, sdt, sc, gc, r):
    """CV_32F guide and / or source through rf_guided_f32 (device entry point and numpy operator)."""
    rng = np.random.default_rng(96)
    n, h, w = 2, 60, 72
    gd = (rng.random((n, h, w, 3)) * 255).astype(np.float32 if gdt == "f" else np.uint8)
    src = (rng.random((n, h, w, 3)) * 200).astype(np.float32 if sdt == "f" else np.uint8)
    gd = gd if gc == 3 else np.ascontiguousarray(gd[..., 0])
    src = src if sc == 3 else np.ascontiguousarray(src[..., 1])
    out = filters.guided_device(torch.from_numpy(gd).cuda(), torch.from_numpy(src).cuda(), r, 3.0).cpu().numpy()
    assert out.dtype == src.dtype and out.shape == src.shape
    for i in range(n):
        ref = oracle.guided(gd[i], src[i], r, 3.0)
        if sdt == "f":
            assert np.abs(out[i] - ref).max() <= 2e-3, (i, np.abs(out[i] - ref).max())
        else:
            mx, frac = lsb_stats(out[i], ref)
            assert mx <= 1 and frac < 2e-3
    got = filters.apply_filter("guided", src[0], gd[0], 3.0, float(r) + 0.5)
    assert got.dtype == src.dtype and np.array_equal(got, out[0].reshape(got.shape))


def test_gf_iterated_equals_repeated_calls():
    """rf_guided_iterated_u8 (guide statistics cached, output fed back through the packed planes) must be
    byte-identical to calling rf_guided_u8 on its own output -- fast path, generic path, gray and colour."""
    for (h, w, r, sc, iters) in [(96, 120, 45, 1, 3), (72, 90, 7, 3, 2), (64, 64, 52, 1, 4), (40, 30, 45, 1, 3),
                                 (50, 70, 70, 3, 3), (130, 700, 20, 1, 3)]:
        n = 3
        gd = np.stack([synth.flat(h, w, 700 + i) for i in range(n)])
        src = np.stack([synth.natural(h, w, 800 + i) for i in range(n)])
        src = src if sc == 3 else np.ascontiguousarray(src[..., 0])
        dg, ds = dev_u8(gd), dev_u8(src)
        want = ds
        for _ in range(iters):
            want = filters.guided_device(dg, want, r, 3.0)
        got = filters.guided_device(dg, ds, r, 3.0, iterations=iters)
        assert torch.equal(got, want), (h, w, r, sc, iters, lsb_stats(got.cpu().numpy(), want.cpu().numpy()))
        # and against the oracle applied repeatedly (first image)
        ref = src[0] if sc == 3 else np.repeat(src[0][..., None], 3, axis=2)
        for _ in range(iters):
            ref = oracle.guided(gd[0], ref, r, 3.0)
        g0 = got[0].cpu().numpy()
        g0 = g0 if sc == 3 else np.repeat(g0[..., None], 3, axis=2)
        mx, frac = lsb_stats(g0, ref)
        assert mx <= 1 and frac < 5e-3, (mx, frac)
    with pytest.raises(ValueError):
        filters.guided_device(dg, ds, r, 3.0, iterations=0)


def test_gf_full_size_three_iterations_and_properties(net):
    # BASELINE config 3 on one image: CNN reflectance, 'flat' guide, c3 s45, 3 iterations with uint8
    # re-quantisation between them
    img = synth.natural(384, 512, 3000)
    gd = synth.flat(384, 512, 3500)
    pipe = pipeline.Pipeline(net)
    out = pipe.cnn_gf(dev_u8(img[None]), dev_u8(gd[None]), 3.0, 45.0, iterations=3).cpu().numpy()[0]
    cur = (cnn.get_reflectance_caffe(net, img) * 255).astype(np.uint8)
    for _ in range(3):
        cur = oracle.guided(gd, cur, 45, 3.0)
    mx, frac = lsb_stats(out, cur)
    assert mx <= 1 and frac < 5e-3, (mx, frac)
    const = np.full((384, 512, 3), 131, np.uint8)
    assert np.array_equal(filters.apply_filter("guided", const, gd, 3.0, 45.0), const)
    g3 = np.repeat(cur[:, :, None], 3, axis=2)                      # replicated src == 1-channel path
    o3 = filters.apply_filter("guided", g3, gd, 3.0, 45.0)
    o1 = filters.apply_filter("guided", cur, gd, 3.0, 45.0)
    assert all(np.array_equal(o3[:, :, c], o1) for c in range(3))


def test_gf_self_guided_suggested_parameters(net):
    """filter_reflectance.py:135-137 suggests `--filter_type=guided --sigma_color=7 --sigma_spatial=52` for filtering
    the CNN prediction WITH ITSELF: guide = the gray reflectance PNG as cv2.imread returns it, three equal channels,
    so cov(I) is singular and only eps*Id keeps the 3x3 inverse finite.  At the suggested eps the cofactor arithmetic
    is still well inside FP32 and the result must be the oracle's, byte for byte up to rounding ties."""
    img = synth.natural(384, 512, 8100)
    r8 = (cnn.get_reflectance_caffe(net, img) * 255).astype(np.uint8)          # <base>-r.png
    g3 = np.repeat(r8[:, :, None], 3, axis=2)                                    # cv2.imread of that PNG
    for eps, ss in ((7.0, 52.0), (3.0, 45.0)):
        out = filters.apply_filter("guided", g3, g3.copy(), eps, ss)
        ref = oracle.guided(g3, g3, int(ss), eps)
        mx, frac = lsb_stats(out, ref)
        assert mx <= 1 and frac < 1e-3, (eps, ss, mx, frac)
        assert np.array_equal(out[:, :, 0], out[:, :, 1]) and np.array_equal(out[:, :, 0], out[:, :, 2])


# ---- seeded random sweep: shapes, radii, channel combinations ------------------------------------------
def test_random_sweep_bf_and_gf_against_the_oracle():
    """48 seeded random cases through the device entry points against the CPU oracle: image sizes from 1 pixel
    to a few strips wide (so strip / tile / row-segment boundaries land everywhere), radii below and above the
    image size and on both sides of the fast-path limits, every joint/src channel combination, explicit d,
    batches (batch == per-image), iterated guided filtering.  +-1 LSB everywhere."""
    rng = np.random.default_rng(20260101)
    worst = {"bf": (0, 0.0), "gf": (0, 0.0)}
    for case in range(24):
        h, w = int(rng.integers(1, 150)), int(rng.integers(1, 200))
        jc, sc = int(rng.choice([1, 3])), int(rng.choice([1, 3]))
        ss = float(rng.choice([0.4, 1.0, 2.7, 5.0, 9.3, 14.0, 22.0]))
        scol = float(rng.choice([3.0, 8.0, 20.0, 60.0]))
        d = int(rng.choice([-1, -1, 5, 12, 31]))
        n = int(rng.integers(1, 4))
        joint = np.stack([synth.natural(h, w, 9000 + 7 * case + i) for i in range(n)])
        src = np.stack([synth.stress(h, w, 9500 + 7 * case + i) for i in range(n)])
        joint = joint if jc == 3 else np.ascontiguousarray(joint[..., 0])
        src = src if sc == 3 else np.ascontiguousarray(src[..., 0])
        if case % 5 == 0:
            joint = src.copy()
            jc = sc
        got = filters.joint_bilateral_device(dev_u8(joint), dev_u8(src), scol, ss, d=d).cpu().numpy()
        for i in range(n):
            ref = oracle.joint_bilateral(joint[i], src[i], d, scol, ss)
            mx, frac = lsb_stats(got[i], ref.reshape(got[i].shape))
            assert mx <= 1 and frac < 2e-3, ("bf", case, (n, h, w, jc, sc, ss, scol, d), mx, frac)
            worst["bf"] = max(worst["bf"], (mx, frac))
    for case in range(24):
        h, w = int(rng.integers(1, 160)), int(rng.integers(1, 900 if case % 6 == 0 else 220))
        sc = int(rng.choice([1, 3]))
        r = int(rng.choice([1, 2, 5, 9, 17, 33, 45, 64, 65, 90]))
        eps = float(rng.choice([0.005, 0.5, 3.0, 7.0, 50.0]))
        iters = int(rng.choice([1, 1, 2, 3]))
        n = int(rng.integers(1, 4))
        guide = np.stack([synth.flat(h, w, 9900 + 5 * case + i) if case % 2 else synth.natural(h, w, 9900 + 5 * case + i)
                          for i in range(n)])
        src = np.stack([synth.natural(h, w, 9950 + 5 * case + i) for i in range(n)])
        src = src if sc == 3 else np.ascontiguousarray(src[..., 0])
        got = filters.guided_device(dev_u8(guide), dev_u8(src), r, eps, iterations=iters).cpu().numpy()
        for i in range(n):
            ref = src[i]
            for _ in range(iters):
                ref = oracle.guided(guide[i], ref, r, eps).reshape(src[i].shape)
            mx, frac = lsb_stats(got[i], ref)
            assert mx <= 1 and frac < 2e-3 * iters, ("gf", case, (n, h, w, sc, r, eps, iters), mx, frac)
            worst["gf"] = max(worst["gf"], (mx, frac))
    print("random sweep worst (max LSB, fraction differing):", worst)


# ---- BASELINE configs 4 and 5 at full image size: oracle on crops ------------------------------------
def _crop_regions(h, w, size):
    """(y0, y1, x0, x1) of the four corners (image borders included), two edge strips and one interior block."""
    m = size
    return [(0, m, 0, m), (0, m, w - m, w), (h - m, h, 0, m), (h - m, h, w - m, w),
            (h // 2, h // 2 + m, 0, m), (0, m, w // 3, w // 3 + m), (h // 2 - 7, h // 2 - 7 + m, w // 2 + 5, w // 2 + 5 + m)]


def _check_on_crops(h, w, halo, size, run_oracle, got, max_frac):
    """A filter output inside a region depends only on the input within `halo` of it: run the oracle on the
    region grown by the halo (clipped at the image border, where the oracle's own border rule applies)."""
    worst = (0, 0.0)
    for (y0, y1, x0, x1) in _crop_regions(h, w, size):
        ya, yb, xa, xb = max(0, y0 - halo), min(h, y1 + halo), max(0, x0 - halo), min(w, x1 + halo)
        ref = run_oracle(ya, yb, xa, xb)[y0 - ya:y1 - ya, x0 - xa:x1 - xa]
        mx, frac = lsb_stats(got[y0:y1, x0:x1], ref)
        assert mx <= 1 and frac <= max_frac, ((y0, y1, x0, x1), mx, frac)
        worst = max(worst, (mx, frac))
    return worst


def test_full_size_config5_4k_bf_and_gf_on_crops(net):
    """3840x2160: CNN -> BF c15 s28 (r=42) and CNN -> GF c3 s45 (flat guide), checked against the oracle on
    corner / edge / interior crops; plus batch-of-2 == single image (no cross-image state)."""
    h, w = 2160, 3840
    img = synth.natural(h, w, 5000)
    gd = synth.flat(h, w, 5500)
    d_img = dev_u8(img[None])
    r8 = pipeline.Pipeline(net).reflectance_u8(d_img)
    r8h = r8.cpu().numpy()[0]
    r8h3 = np.repeat(r8h[:, :, None], 3, axis=2)
    bf = filters.joint_bilateral_device(r8, r8, 15.0, 28.0, gray_replicated=True).cpu().numpy()[0]
    _check_on_crops(h, w, 42, 64, lambda ya, yb, xa, xb: oracle.joint_bilateral(
        r8h3[ya:yb, xa:xb].copy(), r8h3[ya:yb, xa:xb], -1, 15.0, 28.0)[:, :, 0], bf, 2e-3)
    gf = filters.guided_device(dev_u8(gd[None]), r8, 45, 3.0).cpu().numpy()[0]
    _check_on_crops(h, w, 90, 64, lambda ya, yb, xa, xb: oracle.guided(
        gd[ya:yb, xa:xb], r8h3[ya:yb, xa:xb], 45, 3.0)[:, :, 0], gf, 5e-3)
    flipped = r8.flip(1).contiguous()
    two = torch.cat([r8, flipped])
    bf2 = filters.joint_bilateral_device(two, two, 15.0, 28.0, gray_replicated=True)
    assert torch.equal(bf2[0].cpu(), torch.from_numpy(bf))
    assert torch.equal(bf2[1:], filters.joint_bilateral_device(flipped, flipped, 15.0, 28.0, gray_replicated=True))
    # a vertical flip commutes with the filter up to the summation order of the taps: +-1 LSB on rounding ties
    mx, frac = lsb_stats(bf2[1].flip(0).cpu().numpy(), bf)
    assert mx <= 1 and frac < 1e-3, (mx, frac)


def test_full_size_config4_1024x768_cnn_bf_on_crops(net, mlp):
    """1024x768 (IIW scale): CNN (whole image vs the FP32 oracle) -> trunc -> BF(CNN,CNN) c20 s22 on crops."""
    h, w = 768, 1024
    img = synth.natural(h, w, 4000)
    pipe = pipeline.Pipeline(net)
    f32, r8 = net.forward_device(dev_u8(img[None]), want_f32=True, want_u8=True)
    assert np.abs(f32.cpu().numpy()[0] - oracle.mlp_forward(mlp, img)).max() < CNN_TIGHT
    out = pipe.cnn_bf(dev_u8(img[None]), 20.0, 22.0).cpu().numpy()[0]
    r8h3 = np.repeat(r8.cpu().numpy()[0][:, :, None], 3, axis=2)
    _check_on_crops(h, w, 33, 96, lambda ya, yb, xa, xb: oracle.joint_bilateral(
        r8h3[ya:yb, xa:xb].copy(), r8h3[ya:yb, xa:xb], -1, 20.0, 22.0)[:, :, 0], out, 2e-3)


# ---- pipeline, CLI, sharding ---------------------------------------------------------------------
def test_pipeline_cnn_bf_config2(net, mlp):
    img = synth.natural(384, 512, 2000)
    pipe = pipeline.Pipeline(net)
    out = pipe.cnn_bf(dev_u8(img[None]), 20.0, 22.0).cpu().numpy()[0]
    r8 = oracle.quantize_trunc(oracle.mlp_forward(mlp, img))
    r8_gpu = pipe.reflectance_u8(dev_u8(img[None])).cpu().numpy()[0]
    mx, frac = lsb_stats(r8_gpu, r8)
    assert mx <= 1 and frac < 1e-3
    g3 = np.repeat(r8_gpu[:, :, None], 3, axis=2)                   # what cv2.imread makes of <base>-r.png
    ref = oracle.joint_bilateral(g3.copy(), g3, -1, 20, 22)[:, :, 0]
    mx, frac = lsb_stats(out, ref)
    assert mx <= 1 and frac < 1e-3, (mx, frac)


def test_run_host_equals_device_path(net):
    imgs = synth.batch("natural", 5, 48, 64, 8)
    pipe = pipeline.Pipeline(net)
    want = pipe.cnn_bf(dev_u8(imgs), 20.0, 22.0).cpu().numpy()
    pin_in = torch.from_numpy(imgs).pin_memory()
    pin_out = torch.empty((5, 48, 64), dtype=torch.uint8).pin_memory()
    pipe.run_host("cnn_bf", pin_in, pin_out, chunk=2, n_streams=2, sigma_color=20.0, sigma_spatial=22.0)
    assert np.array_equal(pin_out.numpy(), want)
    # sharded == unsharded, independent of shard order
    parts = {}
    for rank in (1, 0):
        lo, hi = pipeline.shard_range(5, rank, 2)
        parts[rank] = pipe.cnn_bf(dev_u8(imgs[lo:hi]), 20.0, 22.0).cpu().numpy()
    assert np.array_equal(np.concatenate([parts[0], parts[1]]), want)
    st = pipeline.aggregate_stats(dev_u8(imgs[..., 0]), dev_u8(want))
    assert st[0] == want.size and st[1] == float(want.astype(np.float64).sum())
    assert st[3] == float(np.abs(want.astype(np.int64) - imgs[..., 0].astype(np.int64)).sum())


def test_cli_end_to_end(tmp_path, net, mlp):
    import cv2
    img = synth.natural(60, 84, 9)
    fin = str(tmp_path / "photo.png")
    cv2.imwrite(fin, img)
    env = dict(os.environ, PYTHONPATH=ROOT)
    subprocess.run([sys.executable, os.path.join(ROOT, "decompose_with_trained_CNN.py"), "--filename_in", fin,
                    "--path_out", str(tmp_path)], check=True, env=env, timeout=600)
    for suffix in ("-r.png", "-r_colorized.png", "-s_colorized.png"):
        assert os.path.exists(str(tmp_path / ("photo" + suffix)))
    r_png = cv2.imread(str(tmp_path / "photo-r.png"), cv2.IMREAD_UNCHANGED)
    assert r_png.ndim == 2
    mx, frac = lsb_stats(r_png, oracle.quantize_trunc(oracle.mlp_forward(mlp, img)))
    assert mx <= 1 and frac < 2e-3
    # colorized outputs follow the reference conventions applied to our reflectance
    gray = cnn.get_reflectance_caffe(net, img)
    refl, shad = iu.colorize(gray, img)
    mx, frac = lsb_stats(cv2.imread(str(tmp_path / "photo-r_colorized.png"), cv2.IMREAD_UNCHANGED),
                         iu.quantize(refl, sRGB=True))
    assert mx <= 1 and frac < 1e-4   # float64 pow on the device vs glibc: identical up to boundary cases
    mx, frac = lsb_stats(cv2.imread(str(tmp_path / "photo-s_colorized.png"), cv2.IMREAD_UNCHANGED),
                         iu.quantize(shad, sRGB=True))
    assert mx <= 1 and frac < 1e-4
    subprocess.run([sys.executable, os.path.join(ROOT, "filter_reflectance.py"), "--filter_type=bilateral",
                    "--sigma_color=20", "--sigma_spatial=22", "--filename_in", str(tmp_path / "photo-r.png"),
                    "--guidance_in", str(tmp_path / "photo-r.png"), "--path_out", str(tmp_path)],
                   check=True, env=env, timeout=600)
    fout = str(tmp_path / "photo-r_bilateral_c20.0s22.0.png")
    assert os.path.exists(fout)
    got = cv2.imread(fout)
    g3 = cv2.imread(str(tmp_path / "photo-r.png"))
    mx, frac = lsb_stats(got, oracle.joint_bilateral(g3.copy(), g3, -1, 20, 22))
    assert mx <= 1 and frac < 2e-3
    subprocess.run([sys.executable, os.path.join(ROOT, "filter_reflectance.py"), "--filter_type=guided",
                    "--sigma_color=3", "--sigma_spatial=45", "--filename_in", str(tmp_path / "photo-r.png"),
                    "--guidance_in", fin, "--path_out", str(tmp_path)], check=True, env=env, timeout=600)
    got = cv2.imread(str(tmp_path / "photo-r_guided_c3.0s45.0.png"))
    mx, frac = lsb_stats(got, oracle.guided(img, g3, 45, 3.0))
    assert mx <= 1 and frac < 2e-3


def test_colorize_device_matches_reference_conventions(G, net):
    # (a) the reference's own colorize + imwrite(sRGB=True) bytes (golden fixture made by importing its image_utils)
    img = G["colorize_image"]
    gray = G["imwrite_gray_in"]
    r8, s8 = cnn.colorize_device(dev_u8(img[None]), torch.from_numpy(gray[None].copy()).cuda())
    for got, want in ((r8, G["colorize_r_png"]), (s8, G["colorize_s_png"])):
        mx, frac = lsb_stats(got.cpu().numpy()[0], want)
        assert mx <= 1 and frac < 2e-3, (mx, frac)
    # (b) a batch, against the host mirror, including an image whose values never exceed 1 (no normalisation)
    imgs = synth.batch("natural", 3, 72, 96, 12)
    imgs[2] = 0
    imgs[2, :4, :4] = 1
    inten = net.forward_device(dev_u8(imgs))[0]
    r8, s8 = cnn.colorize_device(dev_u8(imgs), inten)
    inten_h = inten.cpu().numpy()
    for i in range(3):
        refl, shad = iu.colorize(inten_h[i], imgs[i])
        for got, want in ((r8[i], iu.quantize(refl, sRGB=True)), (s8[i], iu.quantize(shad, sRGB=True))):
            mx, frac = lsb_stats(got.cpu().numpy(), want)
            assert mx <= 1 and frac < 1e-4, (i, mx, frac)


def test_batch_front_end_matches_per_file_cli(tmp_path, net):
    import cv2
    from reflectance_filtering_b200 import batch
    src_dir, out_a, out_b, gdir = (tmp_path / n for n in ("in", "batch", "single", "guide"))
    for d in (src_dir, out_a, out_b, gdir):
        d.mkdir()
    shapes = [(40, 56), (40, 56), (33, 47), (40, 56), (33, 47)]
    for i, (h, w) in enumerate(shapes):
        cv2.imwrite(str(src_dir / ("im%d.png" % i)), synth.natural(h, w, 700 + i))
        cv2.imwrite(str(gdir / ("im%d.png" % i)), synth.flat(h, w, 800 + i))
    (src_dir / "broken.png").write_bytes(b"not a png")
    files = batch.list_inputs(str(src_dir))
    assert len(files) == 6
    # chain CNN -> BF(CNN, CNN), two shards processed in reverse order
    for rank in (1, 0):
        res = batch.run_batch(files, str(out_a), mode="decompose+filter", filter_type="bilateral", sigma_color=20,
                              sigma_spatial=22, chunk=2, rank=rank, world=2, colorized=True)
        assert all("broken" in k for k in res["errors"])
    for i in range(5):
        f = str(src_dir / ("im%d.png" % i))
        cnn.decompose_image(f, str(out_b), net=net)
        filters.read_filter_write("bilateral", str(out_b / ("im%d-r.png" % i)), str(out_b / ("im%d-r.png" % i)),
                                  20.0, 22.0, str(out_b))
        for name in ("im%d-r.png" % i, "im%d-r_bilateral_c20.0s22.0.png" % i, "im%d-r_colorized.png" % i,
                     "im%d-s_colorized.png" % i):
            a = cv2.imread(str(out_a / name), cv2.IMREAD_UNCHANGED)
            b = cv2.imread(str(out_b / name), cv2.IMREAD_UNCHANGED)
            assert a is not None and a.shape == b.shape and np.array_equal(a, b), name
    # filter-only mode with a guidance directory: 3 x GF on colour images
    res = batch.run_batch(files[1:], str(out_a), mode="filter", filter_type="guided", sigma_color=3, sigma_spatial=7,
                          guidance=str(gdir), iterations=3, chunk=4)
    assert not res["errors"] and len(res["written"]) == 5
    img = cv2.imread(str(src_dir / "im2.png"))
    gd = cv2.imread(str(gdir / "im2.png"))
    cur = img
    for _ in range(3):
        cur = filters.apply_filter("guided", cur, gd, 3.0, 7.0)
    assert np.array_equal(cv2.imread(str(out_a / "im2_guided_c3.0s7.0.png")), cur)


def test_batch_front_end_many_same_shape_chunks(tmp_path, net):
    """Seven chunks of one shape: the three rotating pinned output buffers are reused while encoders of
    earlier chunks may still be reading them.  Every file must hold ITS image's bytes."""
    import cv2
    from reflectance_filtering_b200 import batch
    src_dir, out = tmp_path / "in", tmp_path / "out"
    src_dir.mkdir()
    out.mkdir()
    n, h, w = 14, 96, 128
    imgs = [synth.natural(h, w, 900 + i) for i in range(n)]
    for i, im in enumerate(imgs):
        cv2.imwrite(str(src_dir / ("im%02d.png" % i)), im)
    res = batch.run_batch(batch.list_inputs(str(src_dir)), str(out), mode="decompose", chunk=2, io_threads=2)
    assert not res["errors"] and len(res["written"]) == n
    want = net.forward_device(dev_u8(np.stack(imgs)), want_f32=False, want_u8=True)[1].cpu().numpy()
    for i in range(n):
        got = cv2.imread(str(out / ("im%02d-r.png" % i)), cv2.IMREAD_UNCHANGED)
        assert got is not None and np.array_equal(got, want[i]), i


def test_whdr_matches_reference_golden_and_oracle(golden_dir, net):
    """rf_whdr_f32 against the reference's own whdr() (golden_whdr.npz) and the restatement on CNN output."""
    from reflectance_filtering_b200 import whdr
    W = np.load(os.path.join(golden_dir, "golden_whdr.npz"))
    for case in W["cases"]:
        refl, blob, delta = W[case + "_reflectance"], W[case + "_blob"], float(W[case + "_delta"])
        got = whdr.whdr_device(torch.from_numpy(refl).cuda(), torch.from_numpy(blob).cuda(), delta).cpu().numpy()
        np.testing.assert_allclose(got, W[case + "_whdr"], rtol=0, atol=1e-13)   # float64 tree sum vs sequential
        assert got[0] == 0.0
        # the numpy-facing mirror of whdr(reflectance, comparisons, delta) with pixel coordinates
        comps, _ = whdr.get_comparisons_from_blob(blob[1], refl.shape[2], refl.shape[3], delta)
        assert abs(whdr.whdr(refl[1], comps, delta) - W[case + "_whdr"][1]) < 1e-13
    # on the CNN's own output: [n, h, w] float32 straight from forward_device
    imgs = synth.batch("natural", 4, 96, 128, 31)
    r = net.forward_device(dev_u8(imgs), want_f32=True, want_u8=False)[0]
    blob = synth.comparisons(4, 1181, seed=9, min_count=500)
    got = whdr.whdr_device(r, torch.from_numpy(blob).cuda(), 0.1).cpu().numpy()
    rh = r.cpu().numpy()
    want = np.array([oracle.whdr(rh[b][None], blob[b], 0.1) for b in range(4)])
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-13)
    mean, count = whdr.mean_whdr(r, torch.from_numpy(blob).cuda(), 0.1)
    assert count == 4 and abs(mean - want.mean()) < 1e-13
    # error behaviour: coordinate outside the image -> IndexError (numpy's), broken count row -> ValueError
    bad = blob.copy()
    bad[0, 0, 0, 0] = 1.0
    with pytest.raises(IndexError):
        whdr.whdr_device(r, torch.from_numpy(bad).cuda(), 0.1)
    bad = blob.copy()
    bad[2, -1, 0, 0] = np.nan
    with pytest.raises(ValueError):
        whdr.whdr_device(r, torch.from_numpy(bad).cuda(), 0.1)
    with pytest.raises(Exception, match="Expecting 1 or 3 channels"):
        whdr.whdr_device(torch.zeros(1, 2, 4, 4, device="cuda"), torch.from_numpy(blob[:1]).cuda(), 0.1)


def test_native_library_is_the_one_loaded():
    from reflectance_filtering_b200 import _native
    before = _native.launch_count()
    img = synth.natural(16, 16, 1)
    filters.apply_filter("bilateral", img, img, 20, 22)
    assert _native.launch_count() > before
    maps = open("/proc/self/maps").read()
    assert "librf_b200.so" in maps
