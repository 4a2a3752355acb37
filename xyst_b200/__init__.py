"""xyst_b200 -- B200-native (sm_100a) implementation of Xyst's RieCG hot path.

Python here is plumbing only: ctypes bindings of the C ABI (include/xyst_b200.h,
include/xyst_host.h) for tests and bench.py. The compute lives in
libxyst_b200.so (hand-written CUDA) and libxyst_host.so (C++ host mirror of the
reference's solver interface). There is no CPU fallback: loading fails loudly if
the libraries are missing, and every compute call fails without a CUDA device.
"""
from .capi import lib, Context, Params, XystError  # noqa: F401
