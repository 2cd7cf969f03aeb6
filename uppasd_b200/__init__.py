"""uppasd_b200 -- B200-native replacement for UppASD's per-time-step spin-dynamics hot path.

The product is the C-ABI shared library `libuppasd_b200.so` (hand-written sm_100a CUDA, see csrc/ and
include/uppasd_b200.h).  This package only holds the build recipe and thin ctypes bindings that play the
role of the reference's Fortran host in tests and benchmarks.  There is no CPU fallback: importing
`uppasd_b200.capi` without the built library raises.
"""
__all__ = ['build', 'capi', 'host', 'lattice']
