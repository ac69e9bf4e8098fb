"""swiftest_b200 -- B200-native (sm_100a) force-and-drift hot path for Swiftest.

The product is the C-ABI shared library `lib/libswiftest_cuda.so` (include/swiftest_cuda.h) that the
reference's Fortran type-bound procedures call as a drop-in.  This Python package is the
test/benchmark harness over that ABI: `Context` mirrors the reference's array-level interfaces
(swiftest_kick_getacch_int_all, swiftest_drift_all, encounter_check_all_*) with numpy arrays in the Fortran
memory layout (r(3,n) == shape (n,3) C-order).

No CPU fallback exists anywhere in this package.
"""
from ._lib import SwcuError, declared_symbols, load, LIB_PATH  # noqa: F401
from .context import Context, PL, TP, LOOP_TRIANGULAR, LOOP_FLAT, LOOP_AUTO  # noqa: F401
from .shard import partition, tp_block_partition  # noqa: F401
