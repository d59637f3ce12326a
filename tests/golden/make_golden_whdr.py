"""Golden vectors for the WHDR metric, produced by the reference's own code.

Run in the build container (needs /root/reference):  python tests/golden/make_golden_whdr.py
Imports /root/reference/training/layers/whdr_layer.py with a stub ``caffe`` module (the file only needs
``caffe.Layer`` as a base class; ``whdr``, ``_lightness`` and ``get_comparisons_from_blob`` are plain numpy) and
evaluates it on synthetic comparison blobs (synth.comparisons) and synthetic reflectances.  Writes
tests/golden/golden_whdr.npz.  NumPy version matters for one detail (float32 scalar vs Python float comparison,
NEP 50); the version used is stored in the file.
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
from reflectance_filtering_b200 import synth  # noqa: E402


def load_whdr_layer():
    stub = types.ModuleType("caffe")
    stub.Layer = object
    sys.modules.setdefault("caffe", stub)
    spec = importlib.util.spec_from_file_location("ref_whdr_layer", os.path.join(REF, "training", "layers", "whdr_layer.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_whdr_layer()
    out = {"numpy_version": np.array(np.__version__)}
    rng = np.random.default_rng(4242)
    cases = []
    # (name, c, n, h, w, max_comparisons, delta)
    for name, c, n, h, w, m, delta in [("gray", 1, 6, 48, 64, 300, 0.1), ("colour", 3, 5, 40, 56, 200, 0.1),
                                       ("gray_d0", 1, 3, 32, 32, 150, 0.0), ("colour_d25", 3, 3, 24, 40, 1181, 0.25)]:
        refl = rng.uniform(0.0, 1.0, (n, c, h, w)).astype(np.float32)
        # quantise so that many ratios sit exactly on / next to the 1 + delta threshold, and plant tiny values
        refl = (np.round(refl * 40) / 40).astype(np.float32)
        refl[:, :, ::7, ::5] = 1e-9
        blob = synth.comparisons(n, m, seed=rng.integers(1 << 30))
        blob[0, -1, 0, 0] = 0  # first image: no comparisons -> WHDR 0
        per_image = []
        for b in range(n):
            comps, file_name = ref.get_comparisons_from_blob(blob[b], h, w, delta)
            per_image.append(ref.whdr(refl[b], comps, delta))
        out[name + "_reflectance"] = refl
        out[name + "_blob"] = blob
        out[name + "_delta"] = np.array(delta)
        out[name + "_whdr"] = np.array(per_image, np.float64)
        cases.append(name)
        print(name, per_image)
    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, "golden_whdr.npz"), **out)
    print("wrote golden_whdr.npz")


if __name__ == "__main__":
    main()
