"""einsums_b200 — B200-native (sm_100a) device backend for the Einsums contraction hot path.

The product is `lib/libeinsums_b200.so` (C ABI in include/einsums_b200.h) plus the C++20 header
shim in `include/Einsums/` that mirrors the reference's einsum/permute API on DeviceTensor.
This Python package is only the test/bench door onto the same C ABI.
"""
from . import _lib  # noqa: F401
from ._lib import Eb200Error, KF, KF_NAMES, ALG_NAMES, build, lib  # noqa: F401

__all__ = ["_lib", "ops", "Eb200Error", "KF", "KF_NAMES", "ALG_NAMES", "build", "lib"]
