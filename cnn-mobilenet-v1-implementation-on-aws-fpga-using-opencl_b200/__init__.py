"""B200-native MobileNet-V1 1.0-224 inference hot path behind the reference's C host structure.

The product is ``libmnv1.so`` (csrc/, C-ABI in include/mnv1.h).  This Python package is
the thin host-side mirror used by the tests and bench: layer schedule, synthetic data,
ctypes binding.  It never imports anything from ``oracle/``.
"""
from . import layers, synth, binding, shard  # noqa: F401
