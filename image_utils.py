"""Drop-in for the reference's ``image_utils`` module (same six functions); the implementation
lives in ``reflectance-filtering_b200/image_utils.py``."""
from reflectance_filtering_b200.image_utils import (  # noqa: F401
    colorize, imread, imwrite, normalize, quantize, rgb_to_srgb, srgb_lut, srgb_to_rgb)
