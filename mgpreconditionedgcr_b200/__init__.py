"""B200-native MG-preconditioned GCR solve path (drop-in for jing2li/MGPreconditionedGCR's Operator / GCR / MG).

The product is libmgcr_b200.so (hand-written sm_100a CUDA behind the C ABI in include/mgcr_b200.h).  This package
only binds it: `capi` holds the ctypes declarations, `host` the Python mirror of the reference's host classes.
Importing `host` loads the library and raises if it is missing -- there is no CPU path.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
