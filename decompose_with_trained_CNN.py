#!/usr/bin/env python
"""Decompose an image with the direct reflectance prediction CNN on a B200.

Command-line drop-in for the reference's ``decompose_with_trained_CNN.py`` (flags and output
files of /root/reference/decompose_with_trained_CNN.py:133-148); the network runs as an sm_100a
kernel, see ``reflectance-filtering_b200/cnn.py``.

Image convention is always a linear RGB image with shape
channels x height x width in the range 0 - 1.
"""
from __future__ import print_function, division

import argparse
import sys

from reflectance_filtering_b200.cnn import (  # noqa: F401
    Net, caffeBlob_to_imgGrayLinear, decompose_image, get_reflectance_caffe)


def build_parser():
    parser = argparse.ArgumentParser(
        description="""Decompose an image with the direct reflectance
                       prediction CNN.""")
    parser.add_argument("--filename_in",
                        help="Filename of the image which should be decomposed.")
    parser.add_argument("--path_out",
                        help="Where the resulting decompositions should be saved.")
    parser.add_argument("--device", type=int, default=None,
                        help="CUDA device index (additive flag; default: current device)")
    return parser


def main(argv=None):
    parser = build_parser()
    args = parser.parse_args(sys.argv[1:] if argv is None else argv)
    if args.filename_in and args.path_out:
        if args.device is not None:
            from reflectance_filtering_b200 import device as _dev
            _dev.bind_device(args.device)
        decompose_image(args.filename_in, args.path_out)
    else:
        parser.print_help()


if __name__ == "__main__":
    main()
