"""pixelflow_b200 -- B200-native (sm_100a, fp64) implementation of PixelFlow's per-timestep hot path.

The product is libpixelflow_gpu.so (pixelflow_b200/csrc, C ABI in include/pixelflow_gpu.h); this
package is the thin host-side mirror used by tests, bench.py and Python drivers.  Importing the
package does not load the library; the first call does, and fails loudly if it is missing.
"""
from .api import (CASE_NAMES, EXPORTS, FIELDS, LIB_PATH, PixelFlowError, Solver, comm_unique_id,  # noqa: F401
                  load_library, parse_porosity_csv)

__all__ = ["Solver", "load_library", "comm_unique_id", "PixelFlowError", "CASE_NAMES", "FIELDS", "EXPORTS",
           "LIB_PATH", "parse_porosity_csv"]
