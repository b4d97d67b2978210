"""End-to-end path for HOST tensors: pinned host buffers in, pinned host buffer out, the contraction on the device.

This is the step *before* the hot path for callers that keep their tensors on the host (the reference's
host-tensor offload). The output is cut into slabs along C's slowest axis (an uncontracted index),
and three CUDA streams overlap  H2D(slab s+1) | contract(slab s) | D2H(slab s-1)  over a ring of device slab buffers,
so a PCIe-bound call costs ~max(H2D, D2H) instead of H2D + compute + D2H. Operands are uploaded once per call and
sliced on the device (views); nothing is ever permuted.
"""
from __future__ import annotations

import torch

from . import ops


class HostContraction:
    """Reusable plan: device buffers are allocated once and reused by every call."""

    def __init__(self, c_idx, c_shape, a_idx, a_shape, b_idx, b_shape, dtype=torch.float64, slab_bytes=1 << 30, ring=3, device=None):
        self.c_idx, self.a_idx, self.b_idx = list(c_idx), list(a_idx), list(b_idx)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.dtype = dtype
        if not c_shape:
            raise ValueError("HostContraction needs a C of rank >= 1")
        lab = self.c_idx[0]
        if self.c_idx.count(lab) != 1 or self.a_idx.count(lab) + self.b_idx.count(lab) != 1:
            raise ValueError("the slowest C index must be an uncontracted index carried by exactly one operand")
        self.slab_in_a = lab in self.a_idx
        self.slab_axis = (self.a_idx if self.slab_in_a else self.b_idx).index(lab)
        row_bytes = int(torch.tensor(c_shape[1:]).prod().item() if len(c_shape) > 1 else 1) * torch.empty((), dtype=dtype).element_size()
        self.rows = max(1, min(c_shape[0], slab_bytes // max(1, row_bytes)))
        self.c_shape = tuple(c_shape)
        self.A = torch.empty(a_shape, dtype=dtype, device=self.device)
        self.B = torch.empty(b_shape, dtype=dtype, device=self.device)
        self.ring = [torch.empty((self.rows,) + tuple(c_shape[1:]), dtype=dtype, device=self.device) for _ in range(ring)]
        self.s_in, self.s_out = torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def __call__(self, C_out_host, C_in_host, A_host, B_host, alpha=1.0, beta=0.0):
        """C_out_host[...] = beta*C_in_host + alpha * A (x) B. All host tensors pinned; returns after the last D2H finished."""
        for t in (C_out_host, A_host, B_host) + ((C_in_host,) if beta != 0 else ()):
            if not t.is_pinned():
                raise ValueError("host tensors must be pinned (page-locked)")
        s_cmp = torch.cuda.current_stream(self.device)
        ev_start = torch.cuda.Event()
        ev_start.record(s_cmp)
        self.s_in.wait_event(ev_start)
        self.s_out.wait_event(ev_start)
        with torch.cuda.stream(self.s_in):
            self.A.copy_(A_host, non_blocking=True)
            self.B.copy_(B_host, non_blocking=True)
        self.h2d_bytes = A_host.numel() * A_host.element_size() + B_host.numel() * B_host.element_size()
        self.d2h_bytes = 0
        n = self.c_shape[0]
        nslab = (n + self.rows - 1) // self.rows
        free_ev = [None] * len(self.ring)   # D2H of the previous user of this ring slot finished
        for s in range(nslab):
            lo, hi = s * self.rows, min(n, (s + 1) * self.rows)
            buf = self.ring[s % len(self.ring)][: hi - lo]
            with torch.cuda.stream(self.s_in):
                if free_ev[s % len(self.ring)] is not None:
                    self.s_in.wait_event(free_ev[s % len(self.ring)])
                if beta != 0:
                    buf.copy_(C_in_host[lo:hi], non_blocking=True)
                    self.h2d_bytes += buf.numel() * buf.element_size()
                ev_in = torch.cuda.Event()
                ev_in.record(self.s_in)
            s_cmp.wait_event(ev_in)
            if free_ev[s % len(self.ring)] is not None and beta == 0:
                s_cmp.wait_event(free_ev[s % len(self.ring)])
            Av = self.A.narrow(self.slab_axis, lo, hi - lo) if self.slab_in_a else self.A
            Bv = self.B if self.slab_in_a else self.B.narrow(self.slab_axis, lo, hi - lo)
            ops.contract(self.c_idx, buf, self.a_idx, Av, self.b_idx, Bv, alpha, beta, stream=s_cmp)
            ev_cmp = torch.cuda.Event()
            ev_cmp.record(s_cmp)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_cmp)
                C_out_host[lo:hi].copy_(buf, non_blocking=True)
                self.d2h_bytes += buf.numel() * buf.element_size()
                ev_out = torch.cuda.Event()
                ev_out.record(self.s_out)
            free_ev[s % len(self.ring)] = ev_out
        # the caller's stream observes completion of everything
        s_cmp.wait_stream(self.s_out)
        s_cmp.wait_stream(self.s_in)
        return nslab
