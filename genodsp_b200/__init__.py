"""genodsp_b200 -- B200 (sm_100a) implementation of genodsp's per-base operator pipeline.

The product is the CUDA library `genodsp_b200/lib/libgdsp_b200.so` (C-ABI:
include/gdsp_b200.h) plus the C host CLI in `genodsp_b200/host/`.  This Python
package is the thin host-side mirror used by the tests and bench.py.
"""
from . import capi  # noqa: F401
from .capi import GdspError  # noqa: F401


def __getattr__(name):
    if name == "Genome":
        from .genome import Genome
        return Genome
    raise AttributeError(name)
