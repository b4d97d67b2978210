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
