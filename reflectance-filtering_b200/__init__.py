"""B200-native reflectance filtering: the CNN -> joint bilateral / guided
filter hot path of tnestmeyer/reflectance-filtering on sm_100a.

Host-side mirror of the reference's operator surface (same names, argument
meaning and error behaviour); all arithmetic runs in ``csrc/librf_b200.so``
through the C ABI declared in ``include/rf_b200.h``.  There is no CPU
fallback: importing the compute entry points without the built library or
without a CUDA device raises.
"""
__version__ = "0.1.0"
