"""Direct reflectance prediction with the WHDR CNN on the GPU.

Host-side mirror of /root/reference/decompose_with_trained_CNN.py.  :class:`Net` stands where
``caffe.Net(network_file, caffe.TEST, weights=caffemodel)`` stood (:104-106): it reads the
*unchanged* ``network_definition.prototxt`` + ``learned_weights.caffemodel`` (blobs matched by
layer name) and uploads them once per device.  :func:`get_reflectance_caffe` (:82-95) and
:func:`decompose_image` (:98-130) keep the reference's names, arguments, return types, output
file names and error behaviour.  The input transform of ``imgCV2_to_caffeBlob`` (:57-69) is
fused into the kernel as a 256-entry table (``image_utils.srgb_lut``).
"""
from __future__ import division, print_function

import ctypes as C
import os

import numpy as np
import torch

from . import _native, caffe_model, device as dev, image_utils as iu

TEST = 1  # caffe.TEST, accepted and ignored (the deploy graph has no phase-dependent layer)


class Net(object):
    """The deploy network, resident on one CUDA device."""

    def __init__(self, network_file=caffe_model.DEFAULT_PROTOTXT, phase=TEST,
                 weights=caffe_model.DEFAULT_CAFFEMODEL, device=None):
        self.mlp = caffe_model.build_pixel_mlp(network_file, weights)
        if self.mlp.concat != list(range(len(self.mlp.hidden))):
            raise ValueError("the Concat layer must take every hidden map in order; got %r"
                             % (self.mlp.concat,))
        self.device = dev.bind_device(device)
        params = self.mlp.flat_params()
        dims = np.asarray(self.mlp.dims(), np.int32)
        lut = np.ascontiguousarray(iu.srgb_lut(), np.float32)
        handle = C.c_void_p()
        _native.check(_native.lib().rf_cnn_create(
            params.ctypes.data_as(C.c_void_p), dims.ctypes.data_as(C.c_void_p), len(dims) - 1,
            lut.ctypes.data_as(C.c_void_p), C.byref(handle)))
        self._handle = handle
        self.inputs = [self.mlp.input_blob]
        self.outputs = [self.mlp.output_blob]

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                _native.lib().rf_cnn_destroy(h)
            except Exception:
                pass
            self._handle = None

    def forward_device(self, images: torch.Tensor, want_f32: bool = True, want_u8: bool = False,
                       out_f32: torch.Tensor = None, out_u8: torch.Tensor = None):
        """``uint8[N,H,W,3]`` BGR CUDA tensor -> (``float32[N,H,W]`` linear reflectance intensity
        or None, ``uint8[N,H,W]`` = trunc(r * 255) or None); asynchronous on the current stream."""
        dev.check_u8_cuda(images, "images")
        if images.dim() != 4 or images.shape[3] != 3:
            raise ValueError("images must be [N,H,W,3], got %r" % (tuple(images.shape),))
        if images.device != self.device:
            raise ValueError("images live on %s, the network on %s" % (images.device, self.device))
        n, h, w, _ = images.shape
        f32 = u8 = None
        if want_f32:
            f32 = out_f32 if out_f32 is not None else torch.empty((n, h, w), dtype=torch.float32, device=self.device)
        if want_u8:
            u8 = out_u8 if out_u8 is not None else torch.empty((n, h, w), dtype=torch.uint8, device=self.device)
        for t, dt in ((f32, torch.float32), (u8, torch.uint8)):
            if t is not None and (t.dtype != dt or tuple(t.shape) != (n, h, w) or not t.is_contiguous()
                                  or t.device != self.device):
                raise ValueError("output buffer must be a contiguous %s [N,H,W] tensor on %s" % (dt, self.device))
        with torch.cuda.device(self.device):
            dev.bind_device(self.device)
            _native.check(_native.lib().rf_cnn_forward_u8(
                self._handle, dev.ptr(images), n, h, w,
                dev.ptr(f32) if want_f32 else C.c_void_p(0),
                dev.ptr(u8) if want_u8 else C.c_void_p(0), dev.stream_ptr()))
        return f32, u8


_default_nets = {}


def default_net(device=None) -> Net:
    """The shipped model, built once per device (the reference rebuilds caffe.Net per call)."""
    d = dev.bind_device(device)
    net = _default_nets.get(d.index)
    if net is None:
        net = _default_nets[d.index] = Net(device=d)
    return net


def caffeBlob_to_imgGrayLinear(blob):
    """Take a [1,1,H,W] blob and turn it into an H x W image (decompose...py:72-79)."""
    b, c = blob.shape[:2]
    if b != 1 or c != 1:
        raise ValueError("Expecting to get 1 image in mini-batch having 1 channel, "
                         "but got batch size of {} and {} channels".format(b, c))
    return blob[0, 0, :, :]


def get_reflectance_caffe(net, image):
    """Run the image through the network and return the result (decompose...py:82-95):
    ``uint8[H,W,3]`` BGR -> ``float32[H,W]`` linear reflectance intensity."""
    if not isinstance(image, np.ndarray) or image.dtype != np.uint8 or image.ndim != 3 \
            or image.shape[2] != 3:
        raise ValueError("image must be a uint8 H x W x 3 BGR array as cv2.imread returns it")
    dev.bind_device(net.device)
    t = dev.to_device(image, "cnn_in")[None]
    f32, _ = net.forward_device(t, want_f32=True, want_u8=False)
    blob = dev.to_host(f32, "cnn_out")[:, None, :, :]
    return caffeBlob_to_imgGrayLinear(blob)


get_reflectance = get_reflectance_caffe


def percentile_rank(count: int, q: float = 99.9) -> int:
    """0-based rank np.percentile(..., q, method='lower') selects among ``count`` values, with numpy's own
    float64 arithmetic: floor((count - 1) * (q / 100))."""
    return int(np.floor((count - 1) * np.true_divide(np.float64(q), 100)))


def colorize_device(images: torch.Tensor, intensity: torch.Tensor, eps: float = 1e-3):
    """Device version of ``iu.colorize`` followed by ``iu.imwrite(..., sRGB=True)``'s quantisation, for a
    batch: ``uint8[N,H,W,3]`` BGR + ``float32[N,H,W]`` -> (``uint8[N,H,W,3]`` colorized reflectance,
    ``uint8[N,H,W]`` shading) = the bytes of ``<base>-r_colorized.png`` / ``<base>-s_colorized.png``
    (decompose_with_trained_CNN.py:122-128)."""
    dev.check_u8_cuda(images, "images")
    if images.dim() != 4 or images.shape[3] != 3:
        raise ValueError("images must be [N,H,W,3]")
    n, h, w, _ = images.shape
    if intensity.dtype != torch.float32 or tuple(intensity.shape) != (n, h, w) or not intensity.is_cuda \
            or not intensity.is_contiguous():
        raise ValueError("intensity must be a contiguous CUDA float32 [N,H,W] tensor")
    L = _native.lib()
    out_r = torch.empty((n, h, w, 3), dtype=torch.uint8, device=images.device)
    out_s = torch.empty((n, h, w), dtype=torch.uint8, device=images.device)
    with torch.cuda.device(images.device):
        dev.bind_device(images.device)
        ws = torch.empty(int(L.rf_colorize_workspace_bytes(n, h, w)), dtype=torch.uint8, device=images.device)
        _native.check(L.rf_colorize_u8(dev.ptr(images), dev.ptr(intensity), n, h, w, float(eps),
                                       percentile_rank(3 * h * w), percentile_rank(h * w), dev.ptr(out_r),
                                       dev.ptr(out_s), dev.ptr(ws), ws.numel(), dev.stream_ptr()))
    return out_r, out_s


def decompose_image(filename_in, path_out, net=None):
    """Run the intrinsic image decomposition (decompose...py:98-130): writes ``<base>-r.png``
    (linear gray), ``<base>-r_colorized.png`` and ``<base>-s_colorized.png`` (sRGB) into
    ``path_out`` and returns the float32 reflectance intensity."""
    if net is None:
        net = default_net()
    image = iu.imread(filename_in)
    stem = os.path.splitext(os.path.basename(filename_in))[0]
    # one upload; the reflectance, its truncated uint8 form and both colorized outputs are produced on the
    # device with the reference's quantisation points, then written with cv2 exactly as iu.imwrite would
    dev.bind_device(net.device)
    t = dev.to_device(image, "cnn_in")[None]
    f32, u8 = net.forward_device(t, want_f32=True, want_u8=True)
    refl_u8, shad_u8 = colorize_device(t, f32)
    reflectance_gray = dev.to_host(f32, "cnn_out")[0]
    iu.imwrite(os.path.join(path_out, stem + '-r.png'), dev.to_host(u8, "cnn_u8")[0])
    iu.imwrite(os.path.join(path_out, stem + '-r_colorized.png'), dev.to_host(refl_u8, "col_r")[0])
    iu.imwrite(os.path.join(path_out, stem + '-s_colorized.png'), dev.to_host(shad_u8, "col_s")[0])
    return reflectance_gray
