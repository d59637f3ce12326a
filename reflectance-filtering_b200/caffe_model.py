"""Readers for the reference's model artefacts (consumed as data, unchanged).

Replaces what ``caffe.Net(network_file, caffe.TEST, weights=caffemodel)`` does
for this path (/root/reference/decompose_with_trained_CNN.py:104-106): parse
the deploy graph from ``network_definition.prototxt`` and copy the learned
blobs from ``learned_weights.caffemodel`` **by layer name** (Caffe's
``Net::CopyTrainedLayersFrom`` rule); source layers
that carry no blobs or have no namesake in the graph are ignored.

No protobuf runtime or ``caffe_pb2`` is needed: the text format is parsed by a
small tokenizer and the binary file by a varint/length-delimited wire walker
(field numbers per SURVEY.md Appendix B.1).

The graph is compiled into :class:`PixelMlp`, the only topology the hot path
supports: a chain of 1x1 / pad 0 / stride 1 convolutions, each followed by an
in-place ReLU, a channel Concat of (a subset of) those maps, one 1x1
convolution to a single channel and a Sigmoid
(/root/reference/network_definition.prototxt:17-165).  Anything else raises
``ValueError`` -- there is no generic-graph fallback.
"""
from __future__ import annotations

import os
import re
import struct
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_PROTOTXT = os.path.join(_HERE, "model", "network_definition.prototxt")
DEFAULT_CAFFEMODEL = os.path.join(_HERE, "model", "learned_weights.caffemodel")


# --------------------------------------------------------------------------
# prototxt (protobuf text format)
# --------------------------------------------------------------------------
_TOKEN = re.compile(r'\s*(?:(#[^\n]*)|([{}:])|"((?:[^"\\]|\\.)*)"|([^\s{}:"#]+))')


def _tokens(text: str):
    pos = 0
    n = len(text)
    while pos < n:
        m = _TOKEN.match(text, pos)
        if not m:
            if text[pos:].strip() == "":
                return
            raise ValueError("prototxt: cannot tokenize at offset %d" % pos)
        pos = m.end()
        if m.group(1) is not None:
            continue
        if m.group(2) is not None:
            yield ("p", m.group(2))
        elif m.group(3) is not None:
            yield ("s", m.group(3))
        else:
            yield ("w", m.group(4))


def _scalar(kind: str, tok: str):
    if kind == "s":
        return tok
    try:
        return int(tok)
    except ValueError:
        pass
    try:
        return float(tok)
    except ValueError:
        pass
    if tok == "true":
        return True
    if tok == "false":
        return False
    return tok  # enum identifier


def parse_prototxt(text: str) -> Dict[str, list]:
    """Parse protobuf text format into ``{field: [values...]}`` (every field
    is kept as a list because any field may repeat)."""
    toks = list(_tokens(text))
    i = 0

    def message(depth: int) -> Dict[str, list]:
        nonlocal i
        msg: Dict[str, list] = {}
        while i < len(toks):
            kind, tok = toks[i]
            if kind == "p" and tok == "}":
                if depth == 0:
                    raise ValueError("prototxt: unbalanced '}'")
                i += 1
                return msg
            if kind != "w":
                raise ValueError("prototxt: expected a field name, got %r" % (tok,))
            name = tok
            i += 1
            if i < len(toks) and toks[i] == ("p", ":"):
                i += 1
            if i >= len(toks):
                raise ValueError("prototxt: truncated after %r" % name)
            kind, tok = toks[i]
            if kind == "p" and tok == "{":
                i += 1
                msg.setdefault(name, []).append(message(depth + 1))
            else:
                i += 1
                msg.setdefault(name, []).append(_scalar(kind, tok))
        if depth != 0:
            raise ValueError("prototxt: missing '}'")
        return msg

    return message(0)


# --------------------------------------------------------------------------
# caffemodel (protobuf wire format)
# --------------------------------------------------------------------------
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    out = 0
    shift = 0
    while True:
        if pos >= len(buf):
            raise ValueError("caffemodel: truncated varint")
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7
        if shift > 70:
            raise ValueError("caffemodel: varint too long")


def _fields(buf: bytes):
    """Yield ``(field_number, wire_type, value)``; length-delimited values are
    returned as ``bytes`` (memoryview slices of the input)."""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            if pos + ln > n:
                raise ValueError("caffemodel: truncated field %d" % fno)
            val = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            val = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError("caffemodel: unsupported wire type %d" % wt)
        yield fno, wt, val


def _packed_varints(buf: bytes) -> List[int]:
    out = []
    pos = 0
    while pos < len(buf):
        v, pos = _varint(buf, pos)
        out.append(v)
    return out


def _blob(buf: bytes) -> np.ndarray:
    """``BlobProto``: 7 = shape{1: dims}, 5 = float data (packed or repeated),
    1..4 = legacy num/channels/height/width."""
    dims: List[int] = []
    legacy = {}
    chunks: List[np.ndarray] = []
    for fno, wt, val in _fields(buf):
        if fno == 7 and wt == 2:
            for f2, w2, v2 in _fields(val):
                if f2 == 1:
                    dims.extend(_packed_varints(v2) if w2 == 2 else [v2])
        elif fno == 5:
            if wt == 2:
                chunks.append(np.frombuffer(bytes(val), dtype="<f4"))
            elif wt == 5:
                chunks.append(np.frombuffer(bytes(val), dtype="<f4"))
        elif fno in (1, 2, 3, 4) and wt == 0:
            legacy[fno] = val
        elif fno == 8 and wt in (1, 2):
            raise ValueError("caffemodel: double_data blobs are not supported")
    data = np.concatenate(chunks) if chunks else np.zeros(0, np.float32)
    if not dims and legacy:
        dims = [legacy.get(k, 1) for k in (1, 2, 3, 4)]
    if not dims:
        dims = [data.size]
    if int(np.prod(dims)) != data.size:
        raise ValueError("caffemodel: blob shape %r does not match %d values"
                         % (dims, data.size))
    return data.astype(np.float32).reshape(dims)


def read_caffemodel(path: str) -> Dict[str, List[np.ndarray]]:
    """Return ``{layer_name: [blob0, blob1, ...]}`` for every layer that
    carries blobs.  Handles ``NetParameter.layer`` (field 100) and the V1
    ``layers`` list (field 2)."""
    with open(path, "rb") as f:
        buf = f.read()
    out: Dict[str, List[np.ndarray]] = {}
    for fno, wt, val in _fields(buf):
        if wt != 2 or fno not in (100, 2):
            continue
        name_field, blob_field = (1, 7) if fno == 100 else (4, 6)
        name = None
        blobs = []
        for f2, w2, v2 in _fields(val):
            if f2 == name_field and w2 == 2:
                name = bytes(v2).decode("utf-8")
            elif f2 == blob_field and w2 == 2:
                blobs.append(_blob(v2))
        if name is not None and blobs:
            out[name] = blobs
    return out


# --------------------------------------------------------------------------
# graph -> PixelMlp
# --------------------------------------------------------------------------
@dataclass
class PixelMlp:
    """The deploy graph as a per-pixel MLP with skip concatenation.

    ``hidden[i] = (W[out,in], b[out])`` applied as ``h_i = relu(W h_{i-1} + b)``
    with ``h_{-1}`` the 3 linear-RGB inputs; ``fuse = (w[sum of concat widths], b)``
    is applied to the concatenation of ``hidden`` outputs listed in ``concat``
    (indices into ``hidden``; post-ReLU because the ReLUs are in-place, SURVEY
    A.1), followed by a sigmoid.
    """
    hidden: List[Tuple[np.ndarray, np.ndarray]]
    concat: List[int]
    fuse_w: np.ndarray
    fuse_b: float
    input_blob: str = "images"
    output_blob: str = "reflectance_intensity"
    layer_names: List[str] = field(default_factory=list)

    @property
    def n_params(self) -> int:
        return int(sum(w.size + b.size for w, b in self.hidden) + self.fuse_w.size + 1)

    @property
    def macs_per_pixel(self) -> int:
        return int(sum(w.size for w, _ in self.hidden) + self.fuse_w.size)

    def flat_params(self) -> np.ndarray:
        """Parameter block in the layout the CUDA library expects
        (include/rf_b200.h, ``rf_cnn_create``): for each hidden layer W
        row-major ``[out][in]`` then b ``[out]``; then fuse w, then fuse b."""
        parts = []
        for w, b in self.hidden:
            parts += [w.reshape(-1), b.reshape(-1)]
        parts += [self.fuse_w.reshape(-1), np.asarray([self.fuse_b], np.float32)]
        return np.ascontiguousarray(np.concatenate(parts).astype(np.float32))

    def dims(self) -> List[int]:
        return [self.hidden[0][0].shape[1]] + [w.shape[0] for w, _ in self.hidden]


def _one(msg, key, default=None):
    v = msg.get(key)
    if not v:
        return default
    return v[-1]


def _conv_is_1x1(cp) -> bool:
    ks = cp.get("kernel_size", [None])
    kh, kw = _one(cp, "kernel_h"), _one(cp, "kernel_w")
    if kh is not None or kw is not None:
        if (kh or 1) != 1 or (kw or 1) != 1:
            return False
    elif any(k != 1 for k in ks):
        return False
    if any(p != 0 for p in cp.get("pad", [])) or _one(cp, "pad_h", 0) or _one(cp, "pad_w", 0):
        return False
    if any(s != 1 for s in cp.get("stride", [])):
        return False
    if any(d != 1 for d in cp.get("dilation", [])):
        return False
    if _one(cp, "group", 1) != 1:
        return False
    return True


def build_pixel_mlp(prototxt: str = DEFAULT_PROTOTXT,
                    caffemodel: str = DEFAULT_CAFFEMODEL) -> PixelMlp:
    with open(prototxt) as f:
        net = parse_prototxt(f.read())
    blobs = read_caffemodel(caffemodel)
    layers = net.get("layer", [])
    if not layers:
        raise ValueError("prototxt has no 'layer' entries")

    input_blob = None
    produced: Dict[str, int] = {}      # blob name -> index into hidden
    relu_done: Dict[str, bool] = {}
    hidden: List[Tuple[np.ndarray, np.ndarray]] = []
    names: List[str] = []
    concat: List[int] = []
    concat_top = None
    fuse = None
    fuse_top = None
    out_blob = None
    prev_blob = None

    def conv_params(lname, cin):
        if lname not in blobs:
            raise ValueError("caffemodel holds no weights for layer %r" % lname)
        bl = blobs[lname]
        w = bl[0]
        if w.ndim != 4 or w.shape[2:] != (1, 1):
            raise ValueError("layer %r: expected [out,in,1,1] weights, got %r" % (lname, w.shape))
        w = w.reshape(w.shape[0], w.shape[1])
        if w.shape[1] != cin:
            raise ValueError("layer %r: weights take %d channels, graph feeds %d"
                             % (lname, w.shape[1], cin))
        b = bl[1].reshape(-1) if len(bl) > 1 else np.zeros(w.shape[0], np.float32)
        if b.size != w.shape[0]:
            raise ValueError("layer %r: bias size mismatch" % lname)
        return np.ascontiguousarray(w, np.float32), np.ascontiguousarray(b, np.float32)

    for lay in layers:
        ltype = _one(lay, "type")
        lname = _one(lay, "name")
        bottoms = lay.get("bottom", [])
        tops = lay.get("top", [])
        if ltype == "Input":
            input_blob = tops[0]
            shape = _one(_one(lay, "input_param", {}), "shape", {})
            dims = shape.get("dim", [])
            if len(dims) == 4 and dims[1] != 3:
                raise ValueError("input blob must have 3 channels, prototxt says %d" % dims[1])
            prev_blob = input_blob
        elif ltype == "Convolution":
            cp = _one(lay, "convolution_param", {})
            if not _conv_is_1x1(cp):
                raise ValueError("layer %r: only 1x1 / pad 0 / stride 1 convolutions are supported" % lname)
            nout = _one(cp, "num_output")
            src = bottoms[0]
            if concat_top is not None and src == concat_top:
                cin = sum(hidden[i][0].shape[0] for i in concat)
                w, b = conv_params(lname, cin)
                if nout != 1 or w.shape[0] != 1:
                    raise ValueError("layer %r: the fusing convolution must have one output" % lname)
                fuse = (w.reshape(-1), float(b[0]))
                fuse_top = tops[0]
                names.append(lname)
            else:
                if src == input_blob:
                    if hidden:
                        raise ValueError("layer %r: second convolution on the input blob" % lname)
                    cin = 3
                elif src in produced and produced[src] == len(hidden) - 1:
                    if not relu_done.get(src):
                        raise ValueError("layer %r: expected an in-place ReLU on %r first" % (lname, src))
                    cin = hidden[-1][0].shape[0]
                else:
                    raise ValueError("layer %r: bottom %r is not the previous map (only a chain is supported)"
                                     % (lname, src))
                w, b = conv_params(lname, cin)
                if w.shape[0] != nout:
                    raise ValueError("layer %r: num_output %r != weight rows %d" % (lname, nout, w.shape[0]))
                hidden.append((w, b))
                produced[tops[0]] = len(hidden) - 1
                relu_done[tops[0]] = False
                names.append(lname)
        elif ltype == "ReLU":
            if bottoms != tops or bottoms[0] not in produced:
                raise ValueError("layer %r: only in-place ReLU on a convolution output is supported" % lname)
            if _one(_one(lay, "relu_param", {}), "negative_slope", 0) not in (0, 0.0):
                raise ValueError("layer %r: leaky ReLU is not supported" % lname)
            relu_done[bottoms[0]] = True
        elif ltype == "Concat":
            axis = _one(_one(lay, "concat_param", {}), "axis", 1)
            if axis != 1:
                raise ValueError("layer %r: only channel concat is supported" % lname)
            for b in bottoms:
                if b not in produced or not relu_done.get(b):
                    raise ValueError("layer %r: bottom %r is not a ReLU'd convolution map" % (lname, b))
                concat.append(produced[b])
            concat_top = tops[0]
        elif ltype == "Sigmoid":
            if fuse is None or bottoms[0] != fuse_top:
                raise ValueError("layer %r: sigmoid must follow the fusing convolution" % lname)
            out_blob = tops[0]
        else:
            raise ValueError("layer %r: unsupported layer type %r" % (lname, ltype))

    if input_blob is None or not hidden or fuse is None or out_blob is None:
        raise ValueError("prototxt is not an Input -> conv/ReLU chain -> Concat -> conv -> Sigmoid graph")
    if any(not v for v in relu_done.values()):
        raise ValueError("every hidden convolution must be followed by an in-place ReLU")
    return PixelMlp(hidden=hidden, concat=concat, fuse_w=np.ascontiguousarray(fuse[0], np.float32),
                    fuse_b=fuse[1], input_blob=input_blob, output_blob=out_blob, layer_names=names)
