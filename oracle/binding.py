"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so). TEST INFRASTRUCTURE ONLY:
import this from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs — never from
the einsums_b200 package."""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")
MAX_RANK = 12
F32, F64, C64, C128 = 0, 1, 2, 3
ALG_NAMES = ["GENERIC", "DOT", "DIRECT", "GER", "GEMV", "GEMM", "INDETERMINATE"]
_NP2DT = {np.dtype(np.float32): F32, np.dtype(np.float64): F64, np.dtype(np.complex64): C64, np.dtype(np.complex128): C128}


class Tensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("dtype", C.c_int), ("rank", C.c_int), ("index", C.c_int * MAX_RANK),
                ("dim", C.c_int64 * MAX_RANK), ("stride", C.c_int64 * MAX_RANK)]


def build():
    out = subprocess.run(["make", "-C", HERE], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("building the oracle failed:\n" + out.stdout + out.stderr)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        d2 = C.c_double * 2
        _lib.orc_classify_static.argtypes = [C.c_int, C.POINTER(C.c_int)] * 3 + [C.c_int, C.c_int, C.c_int]
        _lib.orc_classify.argtypes = [C.POINTER(Tensor)] * 3 + [C.c_int, C.c_int]
        _lib.orc_einsum_generic.argtypes = [d2, C.POINTER(Tensor), d2, C.POINTER(Tensor), C.POINTER(Tensor), C.c_int, C.c_int, C.c_int]
        _lib.orc_permute.argtypes = [d2, C.POINTER(Tensor), d2, C.POINTER(Tensor), C.POINTER(C.c_int), C.c_int]
        _lib.orc_einsum_f64.argtypes = [C.c_double, C.POINTER(Tensor), C.c_double, C.POINTER(Tensor), C.POINTER(Tensor)]
        _lib.orc_permute_fast_f64.argtypes = [C.c_double, C.POINTER(Tensor), C.c_double, C.POINTER(Tensor), C.POINTER(C.c_int)]
        _lib.orc_blas_load.argtypes = [C.c_char_p]
        _lib.orc_blas_info.restype = C.c_char_p
        _lib.orc_ddot.restype = C.c_double
        vp, i64, dbl = C.c_void_p, C.c_int64, C.c_double
        _lib.orc_dgemm.argtypes = [C.c_char, C.c_char, i64, i64, i64, dbl, vp, i64, vp, i64, dbl, vp, i64]
        _lib.orc_sgemm.argtypes = [C.c_char, C.c_char, i64, i64, i64, C.c_float, vp, i64, vp, i64, C.c_float, vp, i64]
        _lib.orc_zgemm.argtypes = [C.c_char, C.c_char, i64, i64, i64, d2, vp, i64, vp, i64, d2, vp, i64]
        _lib.orc_dgemv.argtypes = [C.c_char, i64, i64, dbl, vp, i64, vp, i64, dbl, vp, i64]
        _lib.orc_ddot.argtypes = [i64, vp, i64, vp, i64]
        _lib.orc_dger.argtypes = [i64, i64, dbl, vp, i64, vp, i64, vp, i64]
        _lib.orc_ddirprod.argtypes = [i64, dbl, vp, i64, vp, i64, vp, i64]
        _lib.orc_blas_set_threads.argtypes = [C.c_int]
    return _lib


def _labels(*packs):
    table = {}
    out = []
    for p in packs:
        out.append([table.setdefault(x, len(table)) for x in p])
    return out


def tensor(arr: np.ndarray, labels) -> Tensor:
    """Describe a numpy array (any strides) for the oracle. Keeps no reference: caller owns arr."""
    t = Tensor()
    t.data = arr.ctypes.data
    t.dtype = _NP2DT[arr.dtype]
    t.rank = arr.ndim
    for i in range(arr.ndim):
        t.index[i] = labels[i]
        t.dim[i] = arr.shape[i]
        t.stride[i] = arr.strides[i] // arr.itemsize
    return t


def classify_static(c_idx, a_idx, b_idx, conj_a=False, conj_b=False, same_types=True) -> str:
    lc, la, lb = _labels(c_idx, a_idx, b_idx)
    arr = lambda v: (C.c_int * max(len(v), 1))(*v)
    r = lib().orc_classify_static(len(lc), arr(lc), len(la), arr(la), len(lb), arr(lb), int(conj_a), int(conj_b), int(same_types))
    return ALG_NAMES[r]


def classify(c_idx, Cn, a_idx, An, b_idx, Bn, conj_a=False, conj_b=False) -> str:
    lc, la, lb = _labels(c_idx, a_idx, b_idx)
    tc, ta, tb = tensor(Cn, lc), tensor(An, la), tensor(Bn, lb)
    return ALG_NAMES[lib().orc_classify(C.byref(tc), C.byref(ta), C.byref(tb), int(conj_a), int(conj_b))]


def _pair(v):
    v = complex(v)
    return (C.c_double * 2)(v.real, v.imag)


def einsum_generic(c_idx, Cn, a_idx, An, b_idx, Bn, alpha=1.0, beta=0.0, conj_a=False, conj_b=False, full_space=True):
    """In-place on Cn (numpy). The reference's host generic algorithm."""
    lc, la, lb = _labels(c_idx, a_idx, b_idx)
    tc, ta, tb = tensor(Cn, lc), tensor(An, la), tensor(Bn, lb)
    rc = lib().orc_einsum_generic(_pair(beta), C.byref(tc), _pair(alpha), C.byref(ta), C.byref(tb), int(conj_a), int(conj_b), int(full_space))
    if rc != 0:
        raise RuntimeError("oracle einsum_generic failed")
    return Cn


def einsum_f64(c_idx, Cn, a_idx, An, b_idx, Bn, alpha=1.0, beta=0.0) -> str:
    """The timed CPU baseline: analyser + BLAS route (OpenBLAS if loaded) / OpenMP generic loops."""
    lc, la, lb = _labels(c_idx, a_idx, b_idx)
    tc, ta, tb = tensor(Cn, lc), tensor(An, la), tensor(Bn, lb)
    r = lib().orc_einsum_f64(float(beta), C.byref(tc), float(alpha), C.byref(ta), C.byref(tb))
    if r < 0:
        raise RuntimeError("oracle einsum_f64 failed")
    return ALG_NAMES[r]


def permute(c_idx, Cn, a_idx, An, alpha=1.0, beta=0.0, conj=False, fast=False):
    perm = [list(a_idx).index(x) for x in c_idx]
    lc, la = _labels(c_idx, a_idx)
    tc, ta = tensor(Cn, lc), tensor(An, la)
    pm = (C.c_int * max(len(perm), 1))(*perm)
    if fast:
        rc = lib().orc_permute_fast_f64(float(beta), C.byref(tc), float(alpha), C.byref(ta), pm)
    else:
        rc = lib().orc_permute(_pair(beta), C.byref(tc), _pair(alpha), C.byref(ta), pm, int(conj))
    if rc != 0:
        raise RuntimeError("oracle permute failed")
    return Cn


def load_openblas() -> str:
    """Find an OpenBLAS inside the Python wheels of this image (numpy/scipy) and hand it to the oracle."""
    import numpy
    cands = []
    base = os.path.dirname(os.path.dirname(numpy.__file__))
    for pat in ("numpy.libs/libscipy_openblas64_*.so", "scipy.libs/libscipy_openblas*.so", "*opencv*libs/libopenblas*.so*"):
        cands += sorted(glob.glob(os.path.join(base, pat)))
    for c in cands:
        if lib().orc_blas_load(c.encode()) == 0:
            return lib().orc_blas_info().decode()
    return "none (naive loops)"
