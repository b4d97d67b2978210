"""ctypes binding of libeinsums_b200.so (the C ABI declared in include/einsums_b200.h).

The library is the product; this module is only the Python door to it, used by the parity tests,
the benchmark and the smoke entry. There is no CPU fallback: if the shared library is missing or
no sm_100 device is usable, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "lib", "libeinsums_b200.so")
CSRC = os.path.join(_HERE, "csrc")

MAX_RANK = 12
MAX_UNIQUE = 24

F32, F64, C64, C128 = 0, 1, 2, 3
ALG_NAMES = ["GENERIC", "DOT", "DIRECT", "GER", "GEMV", "GEMM", "INDETERMINATE"]
KF_NAMES = ["NONE", "TILE_MMA", "TILE_SIMT", "ROWREDUCE", "COLREDUCE", "FULLREDUCE", "ELEMENTWISE", "SCALE_ONLY", "TILE_TCGEN05"]
KF = {n: i for i, n in enumerate(KF_NAMES)}

STATUS_NAMES = {
    0: "OK", 1: "INVALID_ARGUMENT", 2: "DOMAIN", 3: "DIMENSION", 4: "RANK", 5: "UNSUPPORTED",
    6: "OUT_OF_MEMORY", 7: "CUDA", 8: "NO_DEVICE", 9: "NOT_INITIALIZED",
}


class Eb200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"eb200 status {code} ({STATUS_NAMES.get(code, '?')}): {msg}")
        self.code = code


class TensorDesc(C.Structure):
    _fields_ = [
        ("data", C.c_void_p),
        ("dtype", C.c_int32),
        ("rank", C.c_int32),
        ("index", C.c_int32 * MAX_RANK),
        ("stride", C.c_int64 * MAX_RANK),
    ]


class ContractDesc(C.Structure):
    _fields_ = [
        ("n_unique", C.c_int32),
        ("conj_a", C.c_int32),
        ("conj_b", C.c_int32),
        ("extent", C.c_int64 * MAX_UNIQUE),
        ("C", TensorDesc),
        ("A", TensorDesc),
        ("B", TensorDesc),
        ("alpha", C.c_double * 2),
        ("beta", C.c_double * 2),
        ("force_kernel", C.c_int32),
        ("split_k", C.c_int32),
        ("kernel_used", C.c_int32),
        ("launches", C.c_int32),
    ]


def build(verbose: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", CSRC, "-j", str(os.cpu_count() or 4)], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("building libeinsums_b200.so failed:\n" + out.stdout[-4000:] + out.stderr[-4000:])
    if verbose:
        print(out.stdout[-2000:])
    return LIB_PATH


class HostStats(C.Structure):
    _fields_ = [("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("slabs", C.c_int64)]


_lib = None

_REAL = {"s": C.c_float, "d": C.c_double, "c": C.c_float, "z": C.c_double}


def _declare(lib):
    vp, i64, ci, cc = C.c_void_p, C.c_int64, C.c_int, C.c_char
    lib.eb200_last_error_string.restype = C.c_char_p
    lib.eb200_launch_count.restype = C.c_int64
    lib.eb200_launch_count_reset.restype = None
    lib.eb200_last_kernel_name.restype = C.c_char_p
    lib.eb200_set_option.argtypes = [C.c_char_p, ci]
    lib.eb200_get_option.argtypes = [C.c_char_p, C.POINTER(ci)]
    lib.eb200_set_memory_budget.argtypes = [C.c_size_t]
    lib.eb200_memory_stats.argtypes = [C.POINTER(C.c_size_t)] * 4
    lib.eb200_contract.argtypes = [vp, C.POINTER(ContractDesc)]
    lib.eb200_contract_plan.argtypes = [C.POINTER(ContractDesc)]
    lib.eb200_contract_grouped.argtypes = [vp, C.c_int64, C.POINTER(ContractDesc), C.POINTER(C.c_int32)]
    lib.eb200_contract_host.argtypes = [vp, C.POINTER(ContractDesc), C.c_size_t, C.POINTER(HostStats)]
    lib.eb200_malloc.argtypes = [C.POINTER(vp), C.c_size_t]
    lib.eb200_free.argtypes = [vp]
    lib.eb200_malloc_host.argtypes = [C.POINTER(vp), C.c_size_t]
    lib.eb200_free_host.argtypes = [vp]
    for n in ("h2d", "d2h", "d2d", "h2d_async", "d2h_async"):
        getattr(lib, "eb200_memcpy_" + n).argtypes = [vp, vp, C.c_size_t, vp]
    lib.eb200_memset_zero.argtypes = [vp, C.c_size_t, vp]
    lib.eb200_stream_create.argtypes = [C.POINTER(vp)]
    lib.eb200_stream_destroy.argtypes = [vp]
    lib.eb200_stream_synchronize.argtypes = [vp]
    lib.eb200_event_create.argtypes = [C.POINTER(vp)]
    lib.eb200_event_destroy.argtypes = [vp]
    lib.eb200_event_record.argtypes = [vp, vp]
    lib.eb200_event_synchronize.argtypes = [vp]
    lib.eb200_event_elapsed_ms.argtypes = [vp, vp, C.POINTER(C.c_float)]
    lib.eb200_mem_get_info.argtypes = [C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    lib.eb200_ipc_get_handle.argtypes = [vp, vp]
    lib.eb200_ipc_open_handle.argtypes = [vp, C.POINTER(vp)]
    lib.eb200_ipc_close_handle.argtypes = [vp]
    lib.eb200_reduce_slots.argtypes = [vp, ci, ci, vp, i64, i64, C.c_double, vp]
    lib.eb200_copy_strided.argtypes = [vp, ci, ci, C.POINTER(i64), vp, C.POINTER(i64), vp, C.POINTER(i64)]
    for p, rt in _REAL.items():
        sp = C.POINTER(rt)
        getattr(lib, f"eb200_{p}gemm").argtypes = [vp, cc, cc, i64, i64, i64, sp, vp, i64, vp, i64, sp, vp, i64]
        getattr(lib, f"eb200_{p}gemm_strided_batched").argtypes = [vp, cc, cc, i64, i64, i64, sp, vp, i64, i64, vp, i64, i64, sp, vp, i64, i64, i64]
        getattr(lib, f"eb200_{p}gemv").argtypes = [vp, cc, i64, i64, sp, vp, i64, vp, i64, sp, vp, i64]
        getattr(lib, f"eb200_{p}dot").argtypes = [vp, i64, vp, i64, vp, i64, sp]
        getattr(lib, f"eb200_{p}dot_dev").argtypes = [vp, i64, vp, i64, vp, i64, vp]
        getattr(lib, f"eb200_{p}ger").argtypes = [vp, i64, i64, sp, vp, i64, vp, i64, vp, i64]
        getattr(lib, f"eb200_{p}axpy").argtypes = [vp, i64, sp, vp, i64, vp, i64]
        getattr(lib, f"eb200_{p}axpby").argtypes = [vp, i64, sp, vp, i64, sp, vp, i64]
        getattr(lib, f"eb200_{p}scal").argtypes = [vp, i64, sp, vp, i64]
        getattr(lib, f"eb200_{p}copy").argtypes = [vp, i64, vp, i64, vp, i64]
        getattr(lib, f"eb200_{p}fill").argtypes = [vp, i64, sp, vp, i64]
        getattr(lib, f"eb200_{p}dirprod").argtypes = [vp, i64, sp, vp, i64, vp, i64, vp, i64]
        getattr(lib, f"eb200_{p}dirdiv").argtypes = [vp, i64, vp, i64, vp, i64]
        getattr(lib, f"eb200_{p}dirmul").argtypes = [vp, i64, vp, i64, vp, i64]
        getattr(lib, f"eb200_{p}permute").argtypes = [vp, ci, C.POINTER(i64), C.POINTER(i64), C.POINTER(ci), C.POINTER(i64), sp, vp, sp, vp, ci]
    for p, rt in (("c", C.c_float), ("z", C.c_double)):
        sp = C.POINTER(rt)
        getattr(lib, f"eb200_{p}dotc").argtypes = [vp, i64, vp, i64, vp, i64, sp]
        getattr(lib, f"eb200_{p}gerc").argtypes = [vp, i64, i64, sp, vp, i64, vp, i64, vp, i64]
        getattr(lib, f"eb200_{p}lacgv").argtypes = [vp, i64, vp, i64]


def lib():
    """Load (once) and return the ctypes handle. Raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(the backend has no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def check(rc: int):
    if rc != 0:
        raise Eb200Error(rc, lib().eb200_last_error_string().decode(errors="replace"))


def declared_symbols() -> list[str]:
    """Every function name declared in include/einsums_b200.h (macro-expanded by hand here)."""
    import re
    hdr = open(os.path.join(ROOT, "include", "einsums_b200.h")).read()
    names = set(re.findall(r"\b(eb200_[a-z0-9_]+)\s*\(", hdr))
    names = {n for n in names if "##" not in n}
    macro = re.search(r"#define EB200_DECL_BLAS\(P, T\)(.*?)\n\n", hdr, re.S).group(1)
    for stem in re.findall(r"eb200_##P##([a-z_]+)\(", macro):
        for p in "sdcz":
            names.add(f"eb200_{p}{stem}")
    return sorted(names)
