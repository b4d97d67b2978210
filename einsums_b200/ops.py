"""Torch-tensor front door to the C ABI (device pointers + strides in, kernels enqueued on the
current torch stream). PyTorch is plumbing only: memory, streams, torch.distributed.

Index-pack conventions follow the reference's einsum/permute API
(TensorAlgebra/TensorAlgebra.hpp:41-59, TensorAlgebra/Permute.hpp:330-407 in the reference tree):
`contract("abef", C, "mnef", g, "mnab", t, alpha=.25, beta=1)` is
`einsum(1.0, Indices{a,b,e,f}, &C, 0.25, Indices{m,n,e,f}, g, Indices{m,n,a,b}, t)`.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import F32, F64, C64, C128, ContractDesc, KF, KF_NAMES, check, lib

_DTYPES = {torch.float32: F32, torch.float64: F64, torch.complex64: C64, torch.complex128: C128}
_PREFIX = {torch.float32: "s", torch.float64: "d", torch.complex64: "c", torch.complex128: "z"}
_REALC = {torch.float32: C.c_float, torch.float64: C.c_double, torch.complex64: C.c_float, torch.complex128: C.c_double}


def init(device: int = 0):
    check(lib().eb200_init(int(device)))


def _stream(stream=None) -> int:
    if stream is None:
        return torch.cuda.current_stream().cuda_stream
    if isinstance(stream, torch.cuda.Stream):
        return stream.cuda_stream
    return int(stream)


class DeviceView:
    """A dense row-major view of device memory given by a raw pointer -- memory torch did not allocate, e.g. a PEER GPU's buffer opened
    with eb200_ipc_open_handle (torch would attribute such a pointer to the peer's device). Accepted wherever a CUDA tensor is."""
    is_cuda = True

    def __init__(self, ptr: int, shape, dtype):
        self._ptr, self.shape, self.dtype = int(ptr), tuple(int(x) for x in shape), dtype
        st, acc = [], 1
        for e in reversed(self.shape):
            st.append(acc)
            acc *= e
        self._strides = tuple(reversed(st))

    def data_ptr(self):
        return self._ptr

    def dim(self):
        return len(self.shape)

    def stride(self, ax=None):
        return self._strides if ax is None else self._strides[ax]

    def numel(self):
        n = 1
        for e in self.shape:
            n *= e
        return n


def _need_cuda(*ts):
    for t in ts:
        if isinstance(t, DeviceView):
            continue
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise TypeError("einsums_b200 operates on CUDA tensors only (there is no CPU fallback)")


def _scalar(dtype, v):
    """host scalar as the ctypes array the C ABI expects (1 real or (re,im))."""
    rt = _REALC[dtype]
    if dtype.is_complex:
        v = complex(v)
        return (rt * 2)(v.real, v.imag)
    return (rt * 1)(float(v))


def make_desc(c_idx, Ct, a_idx, At, b_idx, Bt, alpha=1.0, beta=0.0, conj_a=False, conj_b=False, force_kernel=0, split_k=0) -> ContractDesc:
    """Build the descriptor exactly as the reference's device launcher does: unique indices in
    order C, A, B (UniqueT<C...,A...,B...>), extents resolved C -> A -> B, per-tensor index tables."""
    unique = []
    for pack in (c_idx, a_idx, b_idx):
        for x in pack:
            if x not in unique:
                unique.append(x)
    if len(unique) > _lib.MAX_UNIQUE:
        raise ValueError("too many distinct indices")
    d = ContractDesc()
    d.n_unique = len(unique)
    for name, pack, t in (("C", c_idx, Ct), ("A", a_idx, At), ("B", b_idx, Bt)):
        if len(pack) != t.dim():
            raise ValueError(f"index pack {pack} does not match tensor rank {t.dim()}")
    # extents resolve C -> A -> B (get_dim_ranges_for_many); every other occurrence of the label, in any tensor and also
    # repeated inside one tensor, must agree -- the C ABI carries no per-tensor shapes, so a mismatch would read or write
    # out of bounds. Same error class as the shim's check_dims (reference: TA/Backends/Dispatch.hpp:125-127).
    extent = {}
    for name, pack, t in (("C", c_idx, Ct), ("A", a_idx, At), ("B", b_idx, Bt)):
        for ax, lab in enumerate(pack):
            n = int(t.shape[ax])
            if extent.setdefault(lab, n) != n:
                raise ValueError(f"dimension_error: index {lab!r} has extent {extent[lab]} but axis {ax} of {name} has {n}")
    for u, lab in enumerate(unique):
        d.extent[u] = extent[lab]
    for td, pack, t in ((d.C, c_idx, Ct), (d.A, a_idx, At), (d.B, b_idx, Bt)):
        td.data = t.data_ptr()
        td.dtype = _DTYPES[t.dtype]
        td.rank = t.dim()
        for ax, lab in enumerate(pack):
            td.index[ax] = unique.index(lab)
            td.stride[ax] = t.stride(ax)
    a, b = complex(alpha), complex(beta)
    d.alpha[0], d.alpha[1] = a.real, a.imag
    d.beta[0], d.beta[1] = b.real, b.imag
    d.conj_a, d.conj_b = int(bool(conj_a)), int(bool(conj_b))
    d.force_kernel = int(force_kernel)
    d.split_k = int(split_k)
    return d


def contract(c_idx, Ct, a_idx, At, b_idx, Bt, alpha=1.0, beta=0.0, conj_a=False, conj_b=False, stream=None, force_kernel=0, split_k=0):
    """C[c_idx] = beta*C + alpha * sum A[a_idx]*B[b_idx] through eb200_contract. Returns the descriptor
    (kernel_used / launches / split_k filled in)."""
    _need_cuda(Ct, At, Bt)
    d = make_desc(c_idx, Ct, a_idx, At, b_idx, Bt, alpha, beta, conj_a, conj_b, force_kernel, split_k)
    check(lib().eb200_contract(_stream(stream), C.byref(d)))
    return d


def contract_host(c_idx, Ct, a_idx, At, b_idx, Bt, alpha=1.0, beta=0.0, conj_a=False, conj_b=False, slab_bytes=0, stream=None):
    """eb200_contract_host: any of C, A, B may be a HOST tensor (pinned for overlap); C is updated in place, the call returns
    when the result is in host memory. Returns the transfer statistics (h2d_bytes, d2h_bytes, slabs)."""
    d = make_desc(c_idx, Ct, a_idx, At, b_idx, Bt, alpha, beta, conj_a, conj_b)
    st = _lib.HostStats()
    check(lib().eb200_contract_host(_stream(stream), C.byref(d), int(slab_bytes), C.byref(st)))
    return st


def make_group(c_idx, Cs, a_idx, As, b_idx, Bs, alpha=1.0, beta=0.0, conj_a=False, conj_b=False):
    """Descriptor array for eb200_contract_grouped (build once, run many times with run_group)."""
    if not (len(Cs) == len(As) == len(Bs)):
        raise ValueError("tensor_compat_error: Block tensors need to have the same number of blocks.")
    n = len(Cs)
    arr = (ContractDesc * max(n, 1))()
    for i in range(n):
        _need_cuda(Cs[i], As[i], Bs[i])
        arr[i] = make_desc(c_idx, Cs[i], a_idx, As[i], b_idx, Bs[i], alpha, beta, conj_a, conj_b)
    return arr, n


def run_group(group, stream=None) -> int:
    """Enqueue a descriptor array from make_group; returns the number of contractions actually dispatched."""
    arr, n = group
    groups = C.c_int32(0)
    check(lib().eb200_contract_grouped(_stream(stream), n, arr, C.byref(groups)))
    return groups.value


def contract_grouped(c_idx, Cs, a_idx, As, b_idx, Bs, alpha=1.0, beta=0.0, conj_a=False, conj_b=False, stream=None):
    """The same contraction over lists of tensors (the blocks of block-diagonal / tiled tensors) through
    eb200_contract_grouped: one stream-ordered group, equal-shape runs with constant pointer steps folded into one launch.
    Returns (descriptors, number of contractions actually dispatched)."""
    group = make_group(c_idx, Cs, a_idx, As, b_idx, Bs, alpha, beta, conj_a, conj_b)
    groups = run_group(group, stream)
    return [group[0][i] for i in range(group[1])], groups


def plan(c_idx, Ct, a_idx, At, b_idx, Bt, **kw) -> str:
    """Kernel family the engine would choose (host-only; works on CPU tensors too)."""
    d = make_desc(c_idx, Ct, a_idx, At, b_idx, Bt, **kw)
    check(lib().eb200_contract_plan(C.byref(d)))
    return KF_NAMES[d.kernel_used]


def permute(c_idx, Ct, a_idx, At, alpha=1.0, beta=0.0, conj=False, stream=None):
    """C[c_idx] = beta*C + alpha*A[a_idx] (same index set, any order)."""
    _need_cuda(Ct, At)
    if sorted(map(str, c_idx)) != sorted(map(str, a_idx)) or len(set(c_idx)) != len(c_idx):
        raise ValueError("permute needs the same distinct indices on both sides")
    r = At.dim()
    perm = [list(a_idx).index(x) for x in c_idx]
    for i in range(r):
        if Ct.shape[i] < At.shape[perm[i]]:
            raise ValueError(f"dimension_error: axis {i} of the output is smaller than axis {perm[i]} of the input")
    i64 = C.c_int64 * max(r, 1)
    dims = i64(*At.shape) if r else i64(0)
    sa = i64(*At.stride()) if r else i64(0)
    sc = i64(*Ct.stride()) if r else i64(0)
    pm = (C.c_int * max(r, 1))(*perm) if r else (C.c_int * 1)(0)
    fn = getattr(lib(), f"eb200_{_PREFIX[At.dtype]}permute")
    check(fn(_stream(stream), r, dims, sa, pm, sc, _scalar(At.dtype, alpha), At.data_ptr(), _scalar(At.dtype, beta), Ct.data_ptr(), int(bool(conj))))


# ---- BLAS-like flat seam on column-major device buffers (pointer + ld), used by the parity tests
def gemm(transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, Cm, ldc, stream=None):
    _need_cuda(A, B, Cm)
    fn = getattr(lib(), f"eb200_{_PREFIX[Cm.dtype]}gemm")
    check(fn(_stream(stream), transa.encode(), transb.encode(), m, n, k, _scalar(Cm.dtype, alpha), A.data_ptr(), lda, B.data_ptr(), ldb,
             _scalar(Cm.dtype, beta), Cm.data_ptr(), ldc))


def gemv(trans, m, n, alpha, A, lda, x, incx, beta, y, incy, stream=None):
    _need_cuda(A, x, y)
    fn = getattr(lib(), f"eb200_{_PREFIX[y.dtype]}gemv")
    check(fn(_stream(stream), trans.encode(), m, n, _scalar(y.dtype, alpha), A.data_ptr(), lda, x.data_ptr(), incx, _scalar(y.dtype, beta),
             y.data_ptr(), incy))


def dot(n, x, incx, y, incy, conj=False, stream=None):
    _need_cuda(x, y)
    p = _PREFIX[x.dtype]
    name = f"eb200_{p}dotc" if (conj and x.dtype.is_complex) else f"eb200_{p}dot"
    out = _scalar(x.dtype, 0)
    check(getattr(lib(), name)(_stream(stream), n, x.data_ptr(), incx, y.data_ptr(), incy, out))
    return complex(out[0], out[1]) if x.dtype.is_complex else out[0]


def ger(m, n, alpha, x, incx, y, incy, A, lda, conj=False, stream=None):
    _need_cuda(x, y, A)
    p = _PREFIX[A.dtype]
    name = f"eb200_{p}gerc" if (conj and A.dtype.is_complex) else f"eb200_{p}ger"
    check(getattr(lib(), name)(_stream(stream), m, n, _scalar(A.dtype, alpha), x.data_ptr(), incx, y.data_ptr(), incy, A.data_ptr(), lda))


def axpy(n, alpha, x, incx, y, incy, stream=None):
    _need_cuda(x, y)
    check(getattr(lib(), f"eb200_{_PREFIX[y.dtype]}axpy")(_stream(stream), n, _scalar(y.dtype, alpha), x.data_ptr(), incx, y.data_ptr(), incy))


def axpby(n, alpha, x, incx, beta, y, incy, stream=None):
    _need_cuda(x, y)
    check(getattr(lib(), f"eb200_{_PREFIX[y.dtype]}axpby")(_stream(stream), n, _scalar(y.dtype, alpha), x.data_ptr(), incx,
                                                            _scalar(y.dtype, beta), y.data_ptr(), incy))


def scal(n, alpha, x, incx, stream=None):
    _need_cuda(x)
    check(getattr(lib(), f"eb200_{_PREFIX[x.dtype]}scal")(_stream(stream), n, _scalar(x.dtype, alpha), x.data_ptr(), incx))


def fill(n, value, x, incx, stream=None):
    _need_cuda(x)
    check(getattr(lib(), f"eb200_{_PREFIX[x.dtype]}fill")(_stream(stream), n, _scalar(x.dtype, value), x.data_ptr(), incx))


def copy(n, x, incx, y, incy, stream=None):
    _need_cuda(x, y)
    check(getattr(lib(), f"eb200_{_PREFIX[y.dtype]}copy")(_stream(stream), n, x.data_ptr(), incx, y.data_ptr(), incy))


def dirprod(n, alpha, x, incx, y, incy, z, incz, stream=None):
    _need_cuda(x, y, z)
    check(getattr(lib(), f"eb200_{_PREFIX[z.dtype]}dirprod")(_stream(stream), n, _scalar(z.dtype, alpha), x.data_ptr(), incx, y.data_ptr(), incy,
                                                              z.data_ptr(), incz))


def launch_count() -> int:
    return int(lib().eb200_launch_count())


def launch_count_reset():
    lib().eb200_launch_count_reset()


def set_option(name: str, value: int):
    """Process-wide tuning / testing knob (eb200_set_option): "tma", "tma_generic", "tma_tile", "tma_dist", "permute_tma",
    "tcgen05_generic", "tcgen05_pair", "scratch_keep_mb" (include/einsums_b200.h documents the values)."""
    check(lib().eb200_set_option(name.encode(), int(value)))


def get_option(name: str) -> int:
    v = C.c_int(0)
    check(lib().eb200_get_option(name.encode(), C.byref(v)))
    return v.value


def last_kernel() -> str:
    """Name of the tile kernel this thread's last contraction launched (eb200_last_kernel_name)."""
    return lib().eb200_last_kernel_name().decode()
