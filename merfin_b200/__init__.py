"""merfin_b200: B200-native per-base k-mer multiplicity evaluation (merfin -hist/-dump/-completeness/-polish scoring).

The product is libmerfin_b200.so (hand-written sm_100a CUDA behind the C ABI in
include/merfin_b200.h).  This package is the thin host-side binding used by the
tests and the benchmark; it never computes on the CPU.
"""
from . import lib  # noqa: F401

__all__ = ["lib", "api", "synth"]
