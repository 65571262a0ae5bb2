"""B200 (sm_100a) target for the kernelgen-perf-tests stencil hot path.

The product is `libb200stencil.so` (csrc/, C ABI in include/b200_stencil.h) plus the C host
drivers under drivers/.  This Python package is a thin ctypes binding used by the tests, by
bench.py and by the multi-process slab engine; it contains no compute of its own and no CPU
fallback: loading fails loudly when the CUDA library has not been built.
"""
from . import capi          # noqa: F401
from .capi import (B200Error, Context, TESTS, TEST_ID, F32, F64, load, sweep, test_info,  # noqa: F401
                   interior_points, kernel_info, launch_count, device_count)
