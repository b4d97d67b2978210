"""Sharding of a contraction along an UNCONTRACTED index across the GPUs of one node.

For `Wabef[a,b,e,f] = beta*W + alpha * sum_mn g[m,n,e,f] t[m,n,a,b]` every slab `a in [a0,a1)` of W depends on the
same slab of t's `a` axis and on all of g: ranks own disjoint contiguous slabs, both operands are replicated
(2 x 2.05 GB at o=40, v=400) and the timed region has NO collective. A collective (all-gather of the slabs, or a
reduce-scatter when a *contracted* index is split) is only needed by consumers that want W replicated.

Pure host logic: testable with the gloo backend on CPU (tests/test_sharding_gloo.py).
"""
from __future__ import annotations

from dataclasses import dataclass


def slab_range(extent: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block distribution: the first `extent % world` ranks get one extra element."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(extent, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


@dataclass
class SlabPlan:
    rank: int
    world: int
    a0: int                 # first `a` owned by this rank
    a1: int                 # one past the last
    sub: list               # [(s0, s1)] sub-slabs processed one after another (each fits `budget_bytes`)
    bytes_per_a: int        # bytes of W per value of `a`
    flops_per_a: float


def plan_slabs(extent_a: int, inner_elems: int, k_elems: int, world: int, rank: int, budget_bytes: int, itemsize: int = 8) -> SlabPlan:
    """Split this rank's slab into sub-slabs whose C block (`rows * inner_elems * itemsize`) fits `budget_bytes`.

    inner_elems: number of C elements per value of `a` (v^3 for Wabef); k_elems: contracted extent (o^2)."""
    a0, a1 = slab_range(extent_a, world, rank)
    bytes_per_a = inner_elems * itemsize
    rows = max(1, min(a1 - a0, budget_bytes // bytes_per_a)) if a1 > a0 else 0
    sub = []
    s = a0
    while s < a1:
        e = min(a1, s + rows)
        sub.append((s, e))
        s = e
    return SlabPlan(rank, world, a0, a1, sub, bytes_per_a, 2.0 * inner_elems * k_elems)


def covers_exactly_once(ranges: list, extent: int) -> bool:
    seen = [0] * extent
    for lo, hi in ranges:
        for i in range(lo, hi):
            seen[i] += 1
    return all(c == 1 for c in seen)


# =====================================================================================================================
# K-split: a CONTRACTED index is distributed. Every rank holds its slice of both operands along that
# index and would produce a full-shape partial C; the partials have to be summed. C stays sharded along its slowest
# axis (rank d owns rows slab_range(E0, world, d)).
#
# Every rank owns a receive buffer with one slot per source rank, exported to its peers with eb200_ipc_get_handle (CUDA IPC).
#   mode "push" (default): ONE local product over all rows (a big GEMM runs near the tensor-pipe peak), then the row slabs go to their
#       owners' slots as peer-to-peer copies over NVLink / NVSwitch on the same stream.
#   mode "p2p": rank s computes the rows of destination d with the C pointer of the contraction aimed AT d's slot s -- the GEMM epilogue
#       itself stores over NVLink tile by tile while the tensor pipe works on the next tile (no staging, no send; the exchange costs
#       1.6 % over the products alone), but `world` products of M / world rows each run further from the peak than one big one.
#   mode "nccl" (no peer access needed): partial rows into local scratch, torch.distributed.reduce to the owner, double buffered.
# After a stream-ordered barrier (a one-element NCCL all-reduce) each rank runs eb200_reduce_slots:
#   C_local = beta * C_local + slot[0] + slot[1] + ... in rank order (bit-reproducible whichever rank finished first).
# Measured, Wmnij o=40 v=400 (8.2e11 flop, C = 20 MB), 8 x B200 (profiles/r2/ksplit_n8.json): one GPU 24.0 ms; push 3.28 ms (7.3x),
# p2p 4.69 ms (5.1x), nccl 6.02 ms (4.0x); 2 GPUs: p2p 13.2 ms = the products alone (13.17).
# =====================================================================================================================
def ksplit_layout(extent0: int, inner_elems: int, world: int) -> dict:
    """Receive-buffer geometry: slot stride = the largest slab (rows * inner_elems), one slot per source rank."""
    rows = [slab_range(extent0, world, d) for d in range(world)]
    max_rows = max(hi - lo for lo, hi in rows) if rows else 0
    return {"rows": rows, "slot_stride": max_rows * inner_elems, "slots": world, "elems": world * max_rows * inner_elems}


class KSplitContraction:
    """C_local[rows of this rank, ...] = beta * C_local + alpha * sum over ALL ranks' slices of the split index.

    c_idx / a_idx / b_idx: index strings; the slowest index of C must be an uncontracted index carried by exactly one operand.
    Every rank passes ITS slices A_shard / B_shard (the split index simply has the local extent). dtype: float64 / float32.
    `contract` and `reduce_slots` are injectable so that the ownership / ordering logic runs under gloo on CPU in the tests
    (the product path always uses the library)."""

    def __init__(self, c_idx, c_shape, a_idx, b_idx, dtype, group=None, mode="push", contract=None, reduce_slots=None, alloc=None, grouped=False):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.c_idx, self.a_idx, self.b_idx = c_idx, a_idx, b_idx
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.c_shape = tuple(int(x) for x in c_shape)
        lead = c_idx[0]
        in_a, in_b = lead in a_idx, lead in b_idx
        if in_a == in_b:
            raise ValueError("K-split needs C's slowest index in exactly one operand")
        self.lead_in_a = in_a
        self.inner = 1
        for x in self.c_shape[1:]:
            self.inner *= x
        self.layout = ksplit_layout(self.c_shape[0], self.inner, self.world)
        self.dtype, self.mode = dtype, mode
        self._contract, self._reduce, self._alloc = contract, reduce_slots, alloc
        self.grouped = grouped
        self.profile = None       # set to a list to collect (phase, CUDA event) pairs of one call (tools/run_ksplit.py)
        self.recv = None          # this rank's receive buffer: [world slots][slot_stride]
        self.peers = None         # views of every rank's receive buffer (p2p mode)
        self._raw, self._opened = None, []
        self._setup()

    # -- buffers ------------------------------------------------------------------------------------------------------
    def _setup(self):
        torch, dist = self.torch, self.dist
        n = max(1, self.layout["elems"])
        self.itemsize = 8 if self.dtype == torch.float64 else 4
        if self.mode in ("p2p", "push"):
            import ctypes as C
            from ._lib import check, lib
            # every step that can fail locally is followed by an agreement among the ranks, so that either ALL of them end up in peer mode
            # or ALL raise (and the caller may fall back to mode "nccl") -- never a mix, which would hang the next collective
            raw, handle, err = C.c_void_p(), C.create_string_buffer(64), None
            try:
                check(lib().eb200_malloc(C.byref(raw), n * self.itemsize))
                self._raw = raw.value
                check(lib().eb200_memset_zero(raw, n * self.itemsize, torch.cuda.current_stream().cuda_stream))
                check(lib().eb200_ipc_get_handle(raw, handle))
            except Exception as e:  # noqa: BLE001
                err = f"{type(e).__name__}: {e}"
            handles = [None] * self.world
            dist.all_gather_object(handles, (err, bytes(handle.raw)), group=self.group)
            bad = [h[0] for h in handles if h[0]]
            if bad:
                self._release_local()
                raise RuntimeError("K-split peer buffers: " + bad[0])
            self.peers = []                                   # raw device pointers of every rank's receive buffer
            try:
                for d in range(self.world):
                    if d == self.rank:
                        self.peers.append(self._raw)
                    else:
                        p = C.c_void_p()
                        check(lib().eb200_ipc_open_handle(C.create_string_buffer(handles[d][1], 64), C.byref(p)))
                        self._opened.append(p.value)
                        self.peers.append(p.value)
            except Exception as e:  # noqa: BLE001
                err = f"{type(e).__name__}: {e}"
            errs = [None] * self.world
            dist.all_gather_object(errs, err, group=self.group)
            bad = [e for e in errs if e]
            if bad:
                for p in self._opened:
                    lib().eb200_ipc_close_handle(p)
                self._opened = []
                dist.barrier(group=self.group)
                self._release_local()
                raise RuntimeError("K-split peer buffers: " + bad[0])
            self.recv_ptr = self._raw
            dev = torch.device("cuda", torch.cuda.current_device())
            if self.mode == "push":
                self.full = torch.empty(self.c_shape, dtype=self.dtype, device=dev)   # the rank's whole partial product
        else:
            alloc = self._alloc or (lambda m: torch.zeros(m, dtype=self.dtype, device="cuda"))
            self.recv = alloc(n)
            self.scratch = [alloc(max(1, self.layout["slot_stride"])) for _ in range(2)]
            self.recv_ptr = self.recv.data_ptr()
            dev = self.recv.device
        self._barrier_buf = torch.zeros(1, dtype=torch.float32, device=dev)

    def _release_local(self):
        if self._raw:
            from ._lib import lib
            lib().eb200_free(self._raw)
            self._raw = None

    def close(self):
        if self.mode in ("p2p", "push") and self._raw:
            from ._lib import lib
            self.torch.cuda.synchronize()
            self.dist.barrier(group=self.group)          # nobody may still be writing into a buffer that goes away
            for p in self._opened:
                lib().eb200_ipc_close_handle(p)
            self.dist.barrier(group=self.group)
            lib().eb200_free(self._raw)
            self._raw, self._opened, self.peers = None, [], None

    # -- one distributed contraction ----------------------------------------------------------------------------------
    def _peer_slot(self, d, src, nrows):
        """rows [0, nrows) of slot `src` in rank d's receive buffer, as something ops.contract accepts as C"""
        from .ops import DeviceView
        return DeviceView(self.peers[d] + src * self.layout["slot_stride"] * self.itemsize, (nrows,) + self.c_shape[1:], self.dtype)

    def _slot(self, buf, src, nrows):
        st = self.layout["slot_stride"]
        return buf[src * st: src * st + nrows * self.inner].view((nrows,) + self.c_shape[1:])

    def __call__(self, C_local, A_shard, B_shard, alpha=1.0, beta=0.0):
        torch, dist = self.torch, self.dist
        contract = self._contract
        if contract is None:
            from . import ops
            contract = ops.contract
        lo_r, hi_r = self.layout["rows"][self.rank]
        if tuple(C_local.shape) != (hi_r - lo_r,) + self.c_shape[1:]:
            raise ValueError(f"dimension_error: C_local must hold this rank's rows {lo_r}:{hi_r}")
        lead_axis = (self.a_idx if self.lead_in_a else self.b_idx).index(self.c_idx[0])
        def mark(name):
            if self.profile is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                self.profile.append((name, ev))

        mark("start")
        works = []
        if self.mode == "push":
            # ONE product over all rows into local memory (a large GEMM runs closer to the tensor-pipe peak than `world` small ones), then the
            # row slabs go to their owners' slots as peer-to-peer copies on this stream
            from ._lib import check, lib
            contract(self.c_idx, self.full, self.a_idx, A_shard, self.b_idx, B_shard, alpha, 0.0)
            st = torch.cuda.current_stream().cuda_stream
            for step in range(self.world):
                d = (self.rank + step) % self.world
                lo, hi = self.layout["rows"][d]
                if hi > lo:
                    check(lib().eb200_memcpy_d2d(self.peers[d] + self.rank * self.layout["slot_stride"] * self.itemsize,
                                                 self.full.data_ptr() + lo * self.inner * self.itemsize, (hi - lo) * self.inner * self.itemsize, st))
        use_group = self.mode == "p2p" and self._contract is None and self.grouped
        if use_group:
            # all destinations in ONE grouped call: the per-destination products (M / world rows each) are too small to fill the GPU one
            # after another; eb200_contract_grouped spreads them over forked streams and joins them back into this one
            from . import ops
            Cs, As, Bs = [], [], []
            for step in range(self.world):
                d = (self.rank + step) % self.world
                lo, hi = self.layout["rows"][d]
                if hi == lo:
                    continue
                Cs.append(self._peer_slot(d, self.rank, hi - lo))
                As.append(A_shard.narrow(lead_axis, lo, hi - lo) if self.lead_in_a else A_shard)
                Bs.append(B_shard if self.lead_in_a else B_shard.narrow(lead_axis, lo, hi - lo))
            if Cs:
                ops.contract_grouped(self.c_idx, Cs, self.a_idx, As, self.b_idx, Bs, alpha, 0.0)
        for step in range(0 if (use_group or self.mode == "push") else self.world):
            d = (self.rank + step) % self.world if self.mode == "p2p" else step   # p2p: spread the NVLink targets; nccl: same order everywhere
            lo, hi = self.layout["rows"][d]
            if hi == lo:
                continue
            A_d = A_shard.narrow(lead_axis, lo, hi - lo) if self.lead_in_a else A_shard
            B_d = B_shard if self.lead_in_a else B_shard.narrow(lead_axis, lo, hi - lo)
            if self.mode == "p2p":
                contract(self.c_idx, self._peer_slot(d, self.rank, hi - lo), self.a_idx, A_d, self.b_idx, B_d, alpha, 0.0)
            else:
                buf = self.scratch[step & 1]
                if len(works) >= 2 and works[-2] is not None:
                    works[-2].wait()                                    # the buffer's previous reduce has read it
                part = buf[: (hi - lo) * self.inner].view((hi - lo,) + self.c_shape[1:])
                contract(self.c_idx, part, self.a_idx, A_d, self.b_idx, B_d, alpha, 0.0)
                # sum over ranks lands in the owner's scratch; the owner then keeps it in slot 0 of its receive buffer
                w = dist.reduce(part, dst=dist.get_global_rank(self.group, d) if self.group is not None else d, group=self.group, async_op=True)
                works.append(w)
                if d == self.rank:
                    w.wait()
                    self._slot(self.recv, 0, hi - lo).copy_(part)
        mark("products")
        if self.mode in ("p2p", "push"):
            dist.all_reduce(self._barrier_buf, group=self.group)        # stream-ordered barrier: every rank's stores into my slots are done
            n_slots = self.world
        else:
            for w in works:
                if w is not None:
                    w.wait()
            n_slots = 1
        mark("barrier")
        count = (hi_r - lo_r) * self.inner
        if count:
            if self._reduce is not None:
                self._reduce(n_slots, self.recv, self.layout["slot_stride"], count, beta, C_local)   # test hook (nccl / gloo mode only)
            else:
                from ._lib import F32, F64, check, lib
                if not C_local.is_contiguous():
                    raise ValueError("C_local must be contiguous")
                check(lib().eb200_reduce_slots(torch.cuda.current_stream().cuda_stream, F64 if self.dtype == torch.float64 else F32, n_slots,
                                               self.recv_ptr, self.layout["slot_stride"], count, float(beta), C_local.data_ptr()))
        mark("slot_sum")
        if self.mode in ("p2p", "push"):
            dist.all_reduce(self._barrier_buf, group=self.group)        # the slots may be overwritten by the next call only after everybody has read them
        mark("barrier2")
        return C_local
