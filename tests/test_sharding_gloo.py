"""world_size-2 (and 3) gloo tests on CPU for the multi-GPU host logic: slab ownership covers the sharded index exactly
once, sub-slabs respect the memory budget, and the max-over-ranks timing reduction bench.py uses behaves."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from einsums_b200.sharding import covers_exactly_once, plan_slabs, slab_range


def test_slab_range_properties():
    for extent in (0, 1, 7, 200, 400, 401):
        for world in (1, 2, 3, 4, 8):
            rs = [slab_range(extent, world, r) for r in range(world)]
            assert covers_exactly_once(rs, extent)
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        slab_range(10, 2, 2)


def test_plan_cfg5_memory_budget():
    o, v = 40, 400
    for world in (1, 2, 4, 8):
        for rank in range(world):
            p = plan_slabs(v, v**3, o * o, world, rank, budget_bytes=60 * 2**30)
            assert p.sub[0][0] == p.a0 and p.sub[-1][1] == p.a1
            assert all((e - s) * p.bytes_per_a <= 60 * 2**30 for s, e in p.sub)
            assert covers_exactly_once(p.sub, v) if world == 1 else True
    assert abs(sum(plan_slabs(v, v**3, o * o, 8, r, 2**40).flops_per_a * (slab_range(v, 8, r)[1] - slab_range(v, 8, r)[0]) for r in range(8)) - 8.192e13) < 1e6


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, extent):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = slab_range(extent, world, rank)
    # every rank "computes" a checksum over its slab; the union must reproduce the serial result
    full = torch.arange(extent, dtype=torch.float64) ** 2
    part = full[lo:hi].sum().reshape(1)
    dist.all_reduce(part, op=dist.ReduceOp.SUM)
    assert abs(part.item() - full.sum().item()) < 1e-9
    ranges = [None] * world
    dist.all_gather_object(ranges, (lo, hi))
    assert covers_exactly_once(ranges, extent)
    # bench.py timing rule: the step time of the job is the MAX over ranks
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == 10.0 + world - 1
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,extent", [(2, 400), (3, 200)])
def test_gloo_slab_ownership(world, extent):
    mp.spawn(_worker, args=(world, _free_port(), extent), nprocs=world, join=True)


# ---- K-split (a contracted index is distributed): ownership, double-buffered reduce order and the final beta step, under gloo -------------
def _ksplit_worker(rank, world, port, e0, inner, k_total):
    from einsums_b200.sharding import KSplitContraction, ksplit_layout
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(77)                      # the same full operands on every rank; each keeps its k-slice
    A = torch.rand(e0, k_total, dtype=torch.float64, generator=g) - 0.5
    B = torch.rand(k_total, inner, dtype=torch.float64, generator=g) - 0.5
    C0 = torch.rand(e0, inner, dtype=torch.float64, generator=g) - 0.5
    k0, k1 = slab_range(k_total, world, rank)
    lo, hi = slab_range(e0, world, rank)

    def cpu_contract(c_idx, C, a_idx, At, b_idx, Bt, alpha, beta):          # stand-in for ops.contract: the TEST's arithmetic, not the product's
        C.copy_(beta * C + alpha * torch.einsum(f"{a_idx},{b_idx}->{c_idx}", At, Bt))

    def cpu_reduce(n_slots, recv, stride, count, beta, C_local):
        acc = beta * C_local.reshape(-1) if beta != 0 else torch.zeros(count, dtype=torch.float64)
        for s in range(n_slots):
            acc = acc + recv[s * stride: s * stride + count]
        C_local.reshape(-1).copy_(acc)

    lay = ksplit_layout(e0, inner, world)
    assert covers_exactly_once(lay["rows"], e0) and lay["slot_stride"] >= (hi - lo) * inner
    ks = KSplitContraction("ij", (e0, inner), "ik", "kj", torch.float64, mode="nccl", contract=cpu_contract, reduce_slots=cpu_reduce,
                           alloc=lambda m: torch.zeros(m, dtype=torch.float64))
    C_local = C0[lo:hi].clone()
    for alpha, beta in ((1.0, 0.0), (0.5, 2.0)):
        ref = beta * C_local + alpha * (A @ B)[lo:hi]
        ks(C_local, A[:, k0:k1].contiguous(), B[k0:k1].contiguous(), alpha, beta)
        assert torch.allclose(C_local, ref, rtol=0, atol=1e-12), (rank, (C_local - ref).abs().max())
    with pytest.raises(ValueError):
        ks(torch.zeros(hi - lo + 1, inner, dtype=torch.float64), A[:, k0:k1], B[k0:k1])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,e0", [(2, 7), (3, 8), (3, 2)])
def test_gloo_ksplit_reduce(world, e0):
    mp.spawn(_ksplit_worker, args=(world, _free_port(), e0, 5, 23), nprocs=world, join=True)


def test_ksplit_rejects_bad_patterns():
    from einsums_b200.sharding import ksplit_layout
    lay = ksplit_layout(10, 6, 4)
    assert lay["rows"] == [(0, 3), (3, 6), (6, 8), (8, 10)] and lay["slot_stride"] == 18 and lay["elems"] == 72
