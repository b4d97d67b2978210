"""CPU tests of the product library's boundary (no GPU, no compute calls): it loads, exports every symbol
include/einsums_b200.h declares, validates arguments with the reference's error classes before touching the
device, plans contractions on the host, and fails LOUDLY (no CPU fallback) when there is no device."""
import ctypes as C

import numpy as np
import pytest
import torch

from einsums_b200 import _lib, ops


def test_exports_every_declared_symbol(eb):
    lib = _lib.lib()
    names = _lib.declared_symbols()
    assert len(names) >= 85
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/einsums_b200.h but not exported: {missing}"
    assert lib.eb200_version() == 1


def test_struct_layout_matches_header(eb):
    # sizeof(eb200_tensor_desc) = 8 + 4 + 4 + 12*4 + 12*8 = 160; contract desc = 4*3 (+4 pad) + 24*8 + 3*160 + 32 + 16
    assert C.sizeof(_lib.TensorDesc) == 160
    assert C.sizeof(_lib.ContractDesc) == 16 + 192 + 480 + 32 + 16


def test_argument_validation_matches_reference_classes(eb):
    """hipBLASVendor/src/gemm.hip:23-80: bad trans -> std::invalid_argument; negative size / small ld -> std::domain_error.
    Validation runs before any CUDA call, so it is testable without a device."""
    lib = _lib.lib()
    one = (C.c_double * 1)(1.0)
    p = C.c_void_p(16)
    assert lib.eb200_dgemm(None, b"x", b"n", 2, 2, 2, one, p, 2, p, 2, one, p, 2) == 1  # INVALID_ARGUMENT
    assert b"transA" in lib.eb200_last_error_string()
    assert lib.eb200_dgemm(None, b"n", b"n", -1, 2, 2, one, p, 2, p, 2, one, p, 2) == 2  # DOMAIN
    assert lib.eb200_dgemm(None, b"n", b"n", 4, 2, 2, one, p, 2, p, 2, one, p, 4) == 2  # lda < m
    assert b"lda" in lib.eb200_last_error_string()
    assert lib.eb200_dgemm(None, b"t", b"n", 4, 2, 3, one, p, 2, p, 3, one, p, 4) == 2  # lda < k for 't'
    # zero sizes return at once (gemm.hip:101-102), even with otherwise invalid arguments
    assert lib.eb200_dgemm(None, b"n", b"n", 0, 2, 2, one, None, 0, None, 0, one, None, 0) == 0
    assert lib.eb200_dgemv(None, b"q", 2, 2, one, p, 2, p, 1, one, p, 1) == 1
    assert lib.eb200_dgemv(None, b"n", 2, 2, one, p, 1, p, 1, one, p, 1) == 2
    # permute: perm must be a permutation
    dims = (C.c_int64 * 2)(2, 3)
    st = (C.c_int64 * 2)(3, 1)
    bad = (C.c_int * 2)(0, 0)
    assert lib.eb200_dpermute(None, 2, dims, st, bad, st, one, p, one, p, 0) == 1


def _plan(c, cs, a, as_, b, bs, dtype=torch.float64, **kw):
    mk = lambda s: torch.empty(s, dtype=dtype)
    return ops.plan(c, mk(cs), a, mk(as_), b, mk(bs), **kw)


def test_planner_families(eb):
    """Host-side planner: which device kernel family each pattern of the benchmark configs lands on."""
    o, v, n = 4, 12, 20
    assert _plan("abef", (v,) * 4, "mnef", (o, o, v, v), "mnab", (o, o, v, v)) == "TILE_MMA"       # Wabef
    assert _plan("mnij", (o,) * 4, "mnef", (o, o, v, v), "ijef", (o, o, v, v)) == "TILE_MMA"       # Wmnij
    assert _plan("mbej", (o, v, v, o), "jnfb", (o, o, v, v), "mnef", (o, o, v, v)) == "TILE_MMA"   # Wmbej, no permute pass
    assert _plan("pq", (n, n), "pqrs", (n,) * 4, "rs", (n, n)) == "ROWREDUCE"                       # Fock J
    assert _plan("pq", (n, n), "prqs", (n,) * 4, "rs", (n, n)) == "ROWREDUCE"                       # Fock K
    assert _plan("q", (n,), "pq", (n, n), "p", (n,)) == "COLREDUCE"                                 # gemv, target contiguous
    assert _plan("", (), "ij", (n, n), "ij", (n, n)) == "FULLREDUCE"                                # dot
    assert _plan("ij", (n, n), "ij", (n, n), "ij", (n, n)) == "ELEMENTWISE"                         # direct
    assert _plan("ab", (n, n), "a", (n,), "b", (n,)) == "ELEMENTWISE"                               # outer
    assert _plan("ij", (n, n), "ik", (n, 0), "kj", (0, n)) == "SCALE_ONLY"                          # empty contraction
    assert _plan("ij", (0, n), "ik", (0, 3), "kj", (3, n)) == "NONE"                                # empty output
    assert _plan("ij", (n, n), "ik", (n, n), "kj", (n, n), dtype=torch.float32) == "TILE_SIMT"
    assert _plan("ij", (n, n), "ik", (n, n), "kj", (n, n), dtype=torch.complex128) == "TILE_SIMT"


def test_no_cpu_fallback(eb):
    """Without a device every compute call must return an error — never a silently computed CPU result."""
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    lib = _lib.lib()
    assert lib.eb200_init(0) in (7, 8)
    with pytest.raises(TypeError):
        ops.contract("ij", torch.zeros(2, 2, dtype=torch.float64), "ik", torch.zeros(2, 2, dtype=torch.float64), "kj",
                     torch.zeros(2, 2, dtype=torch.float64))
    # straight through the C ABI with host pointers posing as device pointers: must fail, must not compute
    a = np.ones((64, 64))
    c = np.zeros((64, 64))
    one, zero = (C.c_double * 1)(1.0), (C.c_double * 1)(0.0)
    rc = lib.eb200_dgemm(None, b"n", b"n", 64, 64, 64, one, a.ctypes.data, 64, a.ctypes.data, 64, zero, c.ctypes.data, 64)
    assert rc != 0 and not c.any()


def test_mixed_dtype_zero_extent_terminates(eb):
    """A mixed-type contraction whose narrower tensor is empty must not recurse (it used to overflow the stack: the empty
    tensor kept its dtype, so the widened descriptor went back into the mixed-type path for ever). Run in a child
    process so that a regression shows up as a failed assertion, not as a dead pytest."""
    import subprocess, sys, textwrap
    code = textwrap.dedent("""
        import ctypes as C, torch
        from einsums_b200 import _lib, ops
        for (cs, as_, bs) in (((4, 4), (4, 0), (0, 4)), ((0, 4), (0, 3), (3, 4))):
            d = ops.make_desc("ij", torch.empty(cs, dtype=torch.float64), "ik", torch.empty(as_, dtype=torch.float32), "kj", torch.empty(bs, dtype=torch.float64))
            d.C.data = d.A.data = d.B.data = 0
            rc = _lib.lib().eb200_contract(None, C.byref(d))
            print("rc", rc)
        print("done")
    """)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, cwd=_lib.ROOT)
    assert out.returncode == 0 and "done" in out.stdout, (out.returncode, out.stdout[-500:], out.stderr[-500:])


def test_make_desc_rejects_inconsistent_extents(eb):
    """The C ABI has no per-tensor shapes: the Python door must refuse a label whose extent differs between tensors (or
    between two axes of one tensor) with the reference's dimension_error class, as the C++ shim's check_dims does."""
    z = lambda *s: torch.empty(s, dtype=torch.float64)
    with pytest.raises(ValueError, match="dimension_error"):
        ops.make_desc("ij", z(5, 3), "ik", z(3, 4), "kj", z(4, 3))
    with pytest.raises(ValueError, match="dimension_error"):
        ops.make_desc("i", z(3), "ii", z(3, 4), "i", z(3))
    with pytest.raises(ValueError, match="dimension_error"):
        ops.make_desc("ij", z(3, 3), "ik", z(3, 4), "kj", z(5, 3))
    ops.make_desc("ij", z(3, 5), "ik", z(3, 4), "kj", z(4, 5))


def test_memory_budget_accounting(eb):
    """Budgeted allocator (reference: GPUMemory/src/ModuleVars.cpp:43-63 try_allocate): a request over the budget is refused
    before any CUDA call, so this part of the contract is testable without a device."""
    lib = _lib.lib()
    lib.eb200_set_memory_budget.argtypes = [C.c_size_t]
    lib.eb200_memory_stats.argtypes = [C.POINTER(C.c_size_t)] * 4
    assert lib.eb200_set_memory_budget(1 << 20) == 0
    p = C.c_void_p()
    assert lib.eb200_malloc(C.byref(p), 2 << 20) == 6  # OUT_OF_MEMORY
    assert b"Could not allocate enough memory on the GPU device" in lib.eb200_last_error_string()
    b, u, pk, n = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_size_t()
    assert lib.eb200_memory_stats(C.byref(b), C.byref(u), C.byref(pk), C.byref(n)) == 0
    assert b.value == 1 << 20 and u.value == 0 and n.value == 0
    assert lib.eb200_set_memory_budget(0) == 0


def test_options_round_trip_without_a_device(eb):
    """eb200_set_option / eb200_get_option: every documented knob exists, keeps its value, and an unknown name is refused (no CUDA call)."""
    defaults = {"tma": 1, "tma_generic": 1, "tma_tile": -1, "tma_dist": 0, "permute_tma": 1, "tcgen05_generic": 1, "tcgen05_pair": 1, "scratch_keep_mb": 4096}
    for name, dflt in defaults.items():
        old = ops.get_option(name)
        assert old == dflt, (name, old)
        ops.set_option(name, 7)
        assert ops.get_option(name) == 7
        ops.set_option(name, old)
    with pytest.raises(Exception):
        ops.set_option("no_such_option", 1)
