"""thingino-accel_b200 -- B200-native drop-in for the thingino-accel `mars` inference hot path.

The product is the C-ABI shared library `lib/libmars_b200.so` (sources in `csrc/`, public
headers in `/include`).  This Python package is the thin host-side mirror used by the tests and
the bench: `capi` binds the library with ctypes (same names and argument meaning as the
reference's include/mars_runtime.h), `marsfile` reads/writes `.mars` files.  Nothing here
computes: without the built CUDA library every call fails loudly.
"""
from . import capi, marsfile, shard  # noqa: F401
from .capi import MarsLibraryMissing, MarsModel, lib  # noqa: F401
