#!/usr/bin/env python
"""Filter a reflectance prediction with a joint bilateral or guided filter on a B200.

Command-line drop-in for the reference's ``filter_reflectance.py`` (flags and output naming of
/root/reference/filter_reflectance.py:99-139); the filters run as sm_100a kernels, see
``reflectance-filtering_b200/filters.py``.
"""
from __future__ import print_function, division

import argparse
import sys

from reflectance_filtering_b200.filters import apply_filter, read_filter_write  # noqa: F401

SUGGESTED = (
    # for filtering the direct CNN prediction with itself
    "--filter_type=bilateral --sigma_color=20 --sigma_spatial=22",
    "--filter_type=guided --sigma_color=7 --sigma_spatial=52",
    # for filtering the direct CNN prediction with 'flat'
    "--filter_type=guided --sigma_color=3 --sigma_spatial=45",
)


def build_parser():
    parser = argparse.ArgumentParser(
        description="""Filter reflectance prediction with a bilateral/guided
                       filter, to enhance piecewise constant reflectance
                       prior.""")
    parser.add_argument("--filename_in",
                        help="Filename of the image which should be filtered.")
    parser.add_argument("--guidance_in",
                        help="Filename of the guidance image which should be used for filtering.")
    parser.add_argument("--path_out",
                        help="Where the resulting decompositions should be saved.")
    parser.add_argument("--sigma_color", type=float, help="color parameter")
    parser.add_argument("--sigma_spatial", type=float, help="spatial parameter")
    parser.add_argument("--filter_type",
                        help="""Which filter to choose, the guided filter (guided) or
                                the joint bilateral filter (bilateral).""")
    parser.add_argument("--device", type=int, default=None,
                        help="CUDA device index (additive flag; default: current device)")
    return parser


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    parser = build_parser()
    args = parser.parse_args(argv)
    if len(argv) > 0:
        if args.device is not None:
            from reflectance_filtering_b200 import device as _dev
            _dev.bind_device(args.device)
        read_filter_write(args.filter_type,
                          args.filename_in, args.guidance_in,
                          args.sigma_color, args.sigma_spatial,
                          args.path_out)
    else:
        parser.print_help()
        print("If you do not have any idea what parameters to choose, " +
              "try one of the following combinations:")
        for line in SUGGESTED:
            print(line)


if __name__ == "__main__":
    main()
