"""Image conventions of the hot path (host side).

Mirror of /root/reference/image_utils.py: same six function names, argument meaning, return
types and error messages, so callers and tests written against the reference read the same.
These are the I/O and colour-space conventions around the GPU kernels -- PNG decode/encode is
host work in the reference too (component 8).  The sRGB->linear transform used on
the CNN path is *not* evaluated here per pixel: :func:`srgb_lut` tabulates it for the 256
possible codes and the table is applied inside the CUDA kernel.

Quirks kept on purpose because they are observable in the outputs:
``rgb_to_srgb`` has the 1.055 factor inside the power (image_utils.py:48), ``imwrite``
quantises by truncation (image_utils.py:68), ``normalize`` only rescales when max > 1 and then
by the 99.9th percentile with 'lower' selection (image_utils.py:87-91).
"""
from __future__ import division, print_function

import cv2
import numpy as np

_SRGB_KNEE = 0.04045
_LINEAR_KNEE = 0.0031308


def srgb_to_rgb(srgb):
    """sRGB -> linear RGB, elementwise (image_utils.py:32-39)."""
    srgb = np.asarray(srgb)
    low = srgb <= _SRGB_KNEE
    with np.errstate(invalid="ignore"):
        curve = np.power((srgb + 0.055) / 1.055, 2.4)
    out = np.where(low, srgb / 12.92, curve)
    # the reference fills np.zeros_like(srgb): the result keeps the input dtype
    return out.astype(srgb.dtype, copy=False)


def rgb_to_srgb(rgb):
    """linear RGB -> "sRGB" as the reference computes it (image_utils.py:42-49): note
    ``(1.055 * x) ** (1 / 2.4) - 0.055``, not the standard ``1.055 * x ** (1 / 2.4) - 0.055``."""
    rgb = np.asarray(rgb)
    low = rgb <= _LINEAR_KNEE
    with np.errstate(invalid="ignore"):
        curve = np.power(1.055 * rgb, 1.0 / 2.4) - 0.055
    out = np.where(low, rgb * 12.92, curve)
    return out.astype(rgb.dtype, copy=False)


def srgb_lut():
    """``float32[256]``: ``srgb_to_rgb(v / 255.0)`` in float64, stored as float32 -- exactly the
    values decompose_with_trained_CNN.py:57-69 + :88 can put into the network's input blob."""
    return srgb_to_rgb(np.arange(256, dtype=np.float64) / 255.0).astype(np.float32)


def imread(filename):
    """cv2.imread with the reference's check (image_utils.py:52-57): always uint8 HxWx3 BGR."""
    img = cv2.imread(filename)
    if img is None:
        raise Exception("Input image not readable: {}".format(filename))
    return img


def normalize(img):
    """image_utils.py:84-92."""
    img = img.copy()
    if np.max(img) > 1:
        img /= np.percentile(img, 99.9, method='lower')
        img = np.clip(img, 0, 1)
    return img


def quantize(image, sRGB=False):
    """The float -> uint8 step of ``imwrite`` (image_utils.py:63-68), separated so device code
    and tests can share it: normalize, optional rgb_to_srgb, ``(x * 255)`` truncated."""
    if image.dtype == np.uint8:
        return image
    image = normalize(image)
    if sRGB:
        image = rgb_to_srgb(image)
    return (image * 255).astype(np.uint8)


def imwrite(filename, image, sRGB=False):
    """image_utils.py:60-73: uint8 images are written untouched, anything else is normalised,
    optionally gamma-encoded and truncated to uint8 first."""
    ok = cv2.imwrite(filename, quantize(image, sRGB=sRGB))
    if not ok:
        raise Exception("Not able to write {}, does the folder exist?".format(filename))


def colorize(intensity, image, eps=1e-3):
    """image_utils.py:76-81: shading = mean_c(image) / intensity ; reflectance = image /
    max(shading, eps).  ``image`` is whatever the caller passes -- decompose_image passes the raw
    0..255 BGR uint8 image (decompose_with_trained_CNN.py:122)."""
    shading = np.mean(image, axis=2) / intensity
    reflectance = image / np.maximum(shading, eps)[:, :, np.newaxis]
    return reflectance, shading
