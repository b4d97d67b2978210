"""einsums_b200 — B200-native (sm_100a) device backend for the Einsums contraction hot path.

The product is `lib/libeinsums_b200.so` (C ABI in include/einsums_b200.h) plus the C++20 header
shim in `include/Einsums/` that mirrors the reference's einsum/permute API on DeviceTensor.
This Python package is the test/bench door onto the same C ABI, plus `compile_plan(...).execute(...)`, the
reference's Python front end (EinsumsPy) for CUDA torch tensors.
"""
from . import _lib  # noqa: F401
from ._lib import Eb200Error, KF, KF_NAMES, ALG_NAMES, build, lib  # noqa: F401

from .plan import (EinsumDirectProductPlan, EinsumDotPlan, EinsumGemmPlan, EinsumGemvPlan, EinsumGenericPlan, EinsumGerPlan,  # noqa: F401
                   compile_plan)

__all__ = ["_lib", "ops", "Eb200Error", "KF", "KF_NAMES", "ALG_NAMES", "build", "lib", "compile_plan", "EinsumGenericPlan", "EinsumDotPlan",
           "EinsumDirectProductPlan", "EinsumGerPlan", "EinsumGemvPlan", "EinsumGemmPlan"]
