"""wass_b200: B200-native dense-stereo stage of WASS (the `wass_stereo` hot path).

The compute lives in hand-written sm_100a CUDA behind the C ABI declared in include/wassgpu.h
(built in-tree as wass_b200/libwassgpu.so).  This Python package is only the ctypes binding used
by tests and bench.py; there is no CPU fallback -- importing `wass_b200.capi` without the built
library raises.
"""
__version__ = "0.1.0"
