"""lambrex_b200 -- B200-native (sm_100a) implementation of LAMBReX's D3Q15 fp64
collide-and-stream hot path.

Layers
  csrc/   CUDA kernels + the C ABI (include/lbx.h)          -> _lib/liblbx.so
  host/   C++17 host mirror of the reference's API (AmrSim) -> _lib/liblambrex.so
  lbx.py  ctypes binding of the C ABI (tests, bench, tools)

There is no CPU fallback: importing the bindings without the built shared
libraries, or initialising them without a CUDA device, raises.
"""
__version__ = "0.1.0"
