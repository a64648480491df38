"""pantax_b200 - B200-native alignment-to-abundance hot path of PanTax.

The product is `libpantax_gpu.so` (hand-written sm_100a CUDA behind the C ABI of
include/pantax_gpu.h).  This package is the thin host-side mirror of the reference's
interface for that path (`pantax_b200.api`) plus the build helper.  There is NO CPU
fallback: importing `api` without the built library, or creating a context without a
CUDA device, raises.
"""
from ._lib import LIB_PATH, PantaxGpuError, load_library  # noqa: F401

__all__ = ["LIB_PATH", "PantaxGpuError", "load_library"]
