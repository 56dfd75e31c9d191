"""hygeia_b200 -- B200-native implementation of Hygeia's inference hot path.

The product is the CUDA library ``csrc/libhygeia_b200.so`` behind the C ABI ``include/hygeia_b200.h``; this package is
the thin host-side mirror of the reference's operator interface.  There is no CPU fallback.
"""
from . import model, philox, synthetic  # noqa: F401

__all__ = ["model", "philox", "synthetic", "single_group"]
