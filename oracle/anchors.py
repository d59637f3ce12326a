"""Independent anchors for the oracle, and the host-CPU reference arm.

TEST INFRASTRUCTURE ONLY (same import rules as ``oracle/__init__.py``).

The reference's arithmetic lives in Caffe and OpenCV-contrib ``ximgproc``; neither is in
this image and neither can be installed offline.  What *is* here and is
third-party code reading the reference's real artefacts / implementing the same
algorithm:

* ``cv2.dnn.readNetFromCaffe(prototxt, caffemodel)`` -- OpenCV's own Caffe importer and
  CPU inference (anchor for decompose_with_trained_CNN.py:82-95).
* ``cv2.bilateralFilter(img, -1, sc, ss)`` -- same algorithm as
  ``jointBilateralFilter_8u`` when joint == src (anchor for filter_reflectance.py:60-64).
* ``cv2.boxFilter(..., normalize=True, borderType=BORDER_REFLECT)`` -- the primitive
  ``ximgproc::guidedFilter`` is built on (filter_reflectance.py:67-70).

These are also what ``bench.py --impl reference`` times.
"""
from __future__ import annotations

import numpy as np

try:  # cv2 is part of the image; keep the import error readable if it is not
    import cv2
except Exception as e:  # pragma: no cover
    cv2 = None
    _cv2_err = e


def _need_cv2():
    if cv2 is None:
        raise RuntimeError("cv2 is required for the oracle anchors: %r" % (_cv2_err,))


def input_blob_reference(image_bgr_u8: np.ndarray) -> np.ndarray:
    """decompose_with_trained_CNN.py:57-69 restated with numpy float64, cast to float32 as
    the assignment into the Caffe blob does (:88)."""
    blob = image_bgr_u8 / 255.0
    blob = blob[:, :, ::-1]
    lin = np.where(blob <= 0.04045, blob / 12.92, np.power((blob + 0.055) / 1.055, 2.4))
    return np.ascontiguousarray(np.transpose(lin, (2, 0, 1))[np.newaxis].astype(np.float32))


class DnnNet:
    """cv2.dnn-backed stand-in for ``caffe.Net(prototxt, caffe.TEST, weights=caffemodel)``."""

    def __init__(self, prototxt: str, caffemodel: str):
        _need_cv2()
        self.net = cv2.dnn.readNetFromCaffe(prototxt, caffemodel)

    def forward(self, image_bgr_u8: np.ndarray) -> np.ndarray:
        self.net.setInput(input_blob_reference(image_bgr_u8))
        out = self.net.forward()
        assert out.shape[0] == 1 and out.shape[1] == 1
        return out[0, 0].copy()

    def layer_params(self):
        """``{layer: (W, b)}`` as cv2.dnn parsed them (independent of our wire reader)."""
        out = {}
        for name in self.net.getLayerNames():
            lid = self.net.getLayerId(name)
            try:
                w = self.net.getParam(lid, 0)
                b = self.net.getParam(lid, 1)
            except cv2.error:
                continue
            if w is not None and w.size:
                out[name] = (np.asarray(w), np.asarray(b))
        return out


def bilateral_self(image_u8: np.ndarray, sigma_color: float, sigma_space: float) -> np.ndarray:
    """cv2.bilateralFilter(d=-1): equals jointBilateralFilter(joint=src copy)."""
    _need_cv2()
    return cv2.bilateralFilter(image_u8, -1, float(sigma_color), float(sigma_space))


def joint_bilateral_numpy(joint_u8: np.ndarray, src_u8: np.ndarray, d: int, sigma_color: float,
                          sigma_space: float, border_type: int = 4) -> np.ndarray:
    """SURVEY A.2 (jointBilateralFilter_8u) written independently of oracle/rf_oracle.c: cv2.copyMakeBorder for the
    REFLECT_101 padding, numpy float32 arithmetic vectorised over the image, taps accumulated in the same raster
    order one at a time.  For joint != src, where cv2.bilateralFilter cannot serve as the anchor."""
    _need_cv2()
    f32 = np.float32
    joint = joint_u8 if joint_u8.ndim == 3 else joint_u8[:, :, None]
    src = src_u8 if src_u8.ndim == 3 else src_u8[:, :, None]
    h, w = src.shape[:2]
    sigma_color = sigma_color if sigma_color > 0 else 1.0
    sigma_space = sigma_space if sigma_space > 0 else 1.0
    radius = max(int(np.rint(sigma_space * 1.5)) if d <= 0 else d // 2, 1)   # cvRound = round-half-even
    cw = np.exp(np.arange(256 * joint.shape[2], dtype=np.float64) ** 2 * (-0.5 / sigma_color ** 2)).astype(f32)
    pad = lambda a: cv2.copyMakeBorder(a, radius, radius, radius, radius, int(border_type), value=0).reshape(
        a.shape[0] + 2 * radius, a.shape[1] + 2 * radius, -1)
    pj, ps = pad(np.ascontiguousarray(joint)).astype(np.int32), pad(np.ascontiguousarray(src)).astype(f32)
    j0 = pj[radius:radius + h, radius:radius + w]
    acc = np.zeros((h, w, src.shape[2]), f32)
    wsum = np.zeros((h, w), f32)
    for i in range(-radius, radius + 1):
        for j in range(-radius, radius + 1):
            if i * i + j * j > radius * radius:
                continue
            sw = f32(np.exp((i * i + j * j) * (-0.5 / sigma_space ** 2)))
            jk = pj[radius + i:radius + i + h, radius + j:radius + j + w]
            alpha = np.abs(j0 - jk).sum(axis=2)
            wt = sw * cw[alpha]
            acc += wt[:, :, None] * ps[radius + i:radius + i + h, radius + j:radius + j + w]
            wsum += wt
    out = np.clip(np.rint(acc / wsum[:, :, None]), 0, 255).astype(np.uint8)
    return out if src_u8.ndim == 3 else out[:, :, 0]


def joint_bilateral_f32_numpy(joint: np.ndarray, src: np.ndarray, d: int, sigma_color: float, sigma_space: float,
                              border_type: int = 4) -> np.ndarray:
    """jointBilateralFilter_32f (SURVEY A.2 last bullet) written independently of oracle/rf_oracle.c: exp table of
    4096 bins per joint channel over the joint image's value range, linear interpolation between bins,
    cv2.copyMakeBorder padding, numpy float32 arithmetic, taps in raster order."""
    _need_cv2()
    f32 = np.float32
    J = joint.reshape(joint.shape[0], joint.shape[1], -1).astype(f32)
    S = src.reshape(src.shape[0], src.shape[1], -1).astype(f32)
    h, w = S.shape[:2]
    jc = J.shape[2]
    radius = max(int(np.rint(sigma_space * 1.5)) if d <= 0 else d // 2, 1)
    nbins = 4096 * jc
    color_range = f32((float(J.max()) - float(J.min())) * jc)
    scale = f32(nbins) / color_range
    lut = np.exp((np.arange(nbins + 2, dtype=np.float64) / float(scale)) ** 2 * (-0.5 / sigma_color ** 2)).astype(f32)
    pad = lambda a: cv2.copyMakeBorder(a, radius, radius, radius, radius, int(border_type), value=0).reshape(
        a.shape[0] + 2 * radius, a.shape[1] + 2 * radius, -1)
    pj, ps = pad(np.ascontiguousarray(J)), pad(np.ascontiguousarray(S))
    j0 = pj[radius:radius + h, radius:radius + w]
    acc = np.zeros((h, w, S.shape[2]), f32)
    wsum = np.zeros((h, w), f32)
    for i in range(-radius, radius + 1):
        for j in range(-radius, radius + 1):
            if i * i + j * j > radius * radius:
                continue
            sw = f32(np.exp((i * i + j * j) * (-0.5 / sigma_space ** 2)))
            jk = pj[radius + i:radius + i + h, radius + j:radius + j + w]
            alpha = np.zeros((h, w), f32)
            for c in range(jc):
                alpha = alpha + np.abs(j0[:, :, c] - jk[:, :, c])
            alpha = alpha * scale
            idx = np.minimum(alpha.astype(np.int32), nbins)
            frac = alpha - idx.astype(f32)
            wt = sw * (lut[idx] + frac * (lut[idx + 1] - lut[idx]))
            acc += wt[:, :, None] * ps[radius + i:radius + i + h, radius + j:radius + j + w]
            wsum += wt
    out = acc / wsum[:, :, None]
    return out.reshape(src.shape)


def box_mean_cv2(plane_f32: np.ndarray, r: int) -> np.ndarray:
    _need_cv2()
    k = 2 * int(r) + 1
    return cv2.boxFilter(plane_f32, cv2.CV_32F, (k, k), normalize=True, borderType=cv2.BORDER_REFLECT)


def _to_src_depth(q: np.ndarray, src: np.ndarray) -> np.ndarray:
    """convertTo(depth of src): uint8 = round-half-even + saturate, float32 = as is"""
    return q.astype(np.float32) if src.dtype == np.float32 else np.clip(np.rint(q), 0, 255).astype(np.uint8)


def guided_cv2box(guide_u8: np.ndarray, src_u8: np.ndarray, radius: int, eps: float) -> np.ndarray:
    """SURVEY A.3 written on cv2.boxFilter with float32 pointwise arithmetic (uint8 or float32 images)."""
    _need_cv2()
    f32 = np.float32
    eps = f32(eps)
    if guide_u8.ndim == 2 or guide_u8.shape[2] == 1:  # 1-channel guide: scalar variance, reciprocal, multiply
        I = guide_u8.reshape(guide_u8.shape[:2]).astype(f32)
        src = src_u8 if src_u8.ndim == 3 else src_u8[:, :, None]
        mean = lambda x: box_mean_cv2(x, radius)
        mI = mean(I)
        inv = f32(1.0) / ((mean(I * I) - mI * mI) + eps)
        out = np.empty(src.shape, src_u8.dtype)
        for si in range(src.shape[2]):
            p = src[:, :, si].astype(f32)
            mp = mean(p)
            al = (mean(p * I) - mp * mI) * inv
            be = mp - al * mI
            q = mean(be) + mean(al) * I
            out[:, :, si] = _to_src_depth(q, src_u8)
        return out if src_u8.ndim == 3 else out[:, :, 0]
    I = [guide_u8[:, :, c].astype(f32) for c in range(3)]
    src = src_u8 if src_u8.ndim == 3 else src_u8[:, :, None]
    mean = lambda x: box_mean_cv2(x, radius)
    mI = [mean(x) for x in I]
    cov = {}
    for k in range(3):
        for l in range(k, 3):
            v = mean(I[k] * I[l]) - mI[k] * mI[l]
            if k == l:
                v = v + eps
            cov[(k, l)] = cov[(l, k)] = v
    cof = {}
    for k in range(3):
        for l in range(k, 3):
            k1, k2, l1, l2 = (k + 1) % 3, (k + 2) % 3, (l + 1) % 3, (l + 2) % 3
            cof[(k, l)] = cof[(l, k)] = cov[(k1, l1)] * cov[(k2, l2)] - cov[(k1, l2)] * cov[(k2, l1)]
    det = cov[(0, 0)] * cof[(0, 0)]
    det = det + cov[(0, 1)] * cof[(0, 1)]
    det = det + cov[(0, 2)] * cof[(0, 2)]
    if eps < 1e-2:
        det = np.where(np.abs(det) < 1e-6, f32(1e-6), det)
    inv = {key: v / det for key, v in cof.items()}
    out = np.empty(src.shape, src_u8.dtype)
    for si in range(src.shape[2]):
        p = src[:, :, si].astype(f32)
        mp = mean(p)
        c = [mean(p * I[g]) - mp * mI[g] for g in range(3)]
        al = []
        for g in range(3):
            v = inv[(g, 0)] * c[0]
            v = v + inv[(g, 1)] * c[1]
            v = v + inv[(g, 2)] * c[2]
            al.append(v)
        be = mp
        for g in range(3):
            be = be - al[g] * mI[g]
        mb = mean(be)
        ma = [mean(a) for a in al]
        q = mb
        for g in range(3):
            q = q + ma[g] * I[g]
        out[:, :, si] = _to_src_depth(q, src_u8)
    return out if src_u8.ndim == 3 else out[:, :, 0]


def _box_mean_f64(x: np.ndarray, r: int) -> np.ndarray:
    """Mean over the (2r+1)^2 window with symmetric padding (edge pixel repeated: fedcba|abcdefgh|hgfedcb), float64,
    by a summed-area table.  x: [H, W] or [H, W, K]."""
    pad = [(r + 1, r), (r + 1, r)] + [(0, 0)] * (x.ndim - 2)
    xp = np.pad(x.astype(np.float64), pad, mode="symmetric")
    sat = xp.cumsum(axis=0).cumsum(axis=1)
    k = 2 * r + 1
    h, w = x.shape[:2]
    tot = sat[k:k + h, k:k + w] - sat[:h, k:k + w] - sat[k:k + h, :w] + sat[:h, :w]
    return tot / float(k * k)


def guided_float64(guide_u8: np.ndarray, src_u8: np.ndarray, radius: int, eps: float) -> np.ndarray:
    """The guided filter as the paper states it (He, Sun, Tang, "Guided Image Filtering", eqs. 19-21 of the colour
    case), entirely in float64: a_k = (Sigma_k + eps U)^-1 (mean(I p) - mu_k mean(p)) by np.linalg.solve,
    b_k = mean(p) - a_k . mu_k, q = mean(a) . I + mean(b).  No OpenCV code path, no float32, no cofactors: a third,
    structurally different formulation used as a BOUND on the restatements (they follow guided_filter.cpp's float32
    arithmetic and may differ from this by rounding, not by more)."""
    r = int(radius)
    G = guide_u8.reshape(guide_u8.shape[0], guide_u8.shape[1], -1).astype(np.float64)
    S = src_u8.reshape(src_u8.shape[0], src_u8.shape[1], -1).astype(np.float64)
    gc = G.shape[2]
    mu = _box_mean_f64(G, r)                                                   # [H, W, gc]
    corr = _box_mean_f64(G[:, :, :, None] * G[:, :, None, :], r) if gc > 1 else None
    if gc > 1:
        corr = corr.reshape(G.shape[0], G.shape[1], gc, gc)
        sigma = corr - mu[:, :, :, None] * mu[:, :, None, :] + eps * np.eye(gc)
    else:
        sigma = (_box_mean_f64(G[:, :, 0] ** 2, r) - mu[:, :, 0] ** 2 + eps)[:, :, None, None]
    out = np.empty(S.shape, src_u8.dtype)
    for c in range(S.shape[2]):
        p = S[:, :, c]
        mp = _box_mean_f64(p, r)
        cov_ip = _box_mean_f64(G * p[:, :, None], r) - mu * mp[:, :, None]     # [H, W, gc]
        a = np.linalg.solve(sigma, cov_ip[:, :, :, None])[:, :, :, 0]
        b = mp - (a * mu).sum(axis=2)
        q = (_box_mean_f64(a, r) * G).sum(axis=2) + _box_mean_f64(b, r)
        out[:, :, c] = _to_src_depth(q, src_u8)
    return out.reshape(src_u8.shape)


def quantize_like_imwrite(gray_f32: np.ndarray) -> np.ndarray:
    """image_utils.py:60-73 for the CNN output: normalize is a no-op (sigmoid < 1), then
    ``(image * 255).astype(np.uint8)``; cv2.imread of the 1-channel PNG replicates it to
    three equal channels."""
    return (gray_f32 * 255).astype(np.uint8)
