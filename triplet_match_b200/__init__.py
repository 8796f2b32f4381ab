"""B200-native search path of richard-vock/triplet_match (sm_100a CUDA behind a C-ABI).

The product is the shared library built from csrc/ (C-ABI in include/tm_b200.h)
and the C++ drop-in headers in include/triplet_match/.  The Python modules here
are harness glue only: `capi` (ctypes mirror of the C-ABI) and `synth`
(seed-fixed synthetic clouds and recorded sample lists).
"""
__version__ = "0.1.0"
