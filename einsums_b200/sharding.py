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
