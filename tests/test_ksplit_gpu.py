"""K-split sharding on real GPUs (needs >= 2 devices; skipped on a single-GPU box): all three exchange modes against the full-K product
computed on one GPU, with ragged row / k partitions, beta != 0 and repeated calls (the receive slots are reused)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port):
    import torch.distributed as dist
    from einsums_b200 import ops
    from einsums_b200.sharding import KSplitContraction, slab_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    ops.init(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    o, v = 7, 30                                                      # 7 rows over 2 ranks: 4 + 3; e = 30 over 2 ranks
    gen = torch.Generator(device="cuda").manual_seed(5)
    g = torch.rand(o, o, v, v, dtype=torch.float64, device="cuda", generator=gen) - 0.5
    t = torch.rand(o, o, v, v, dtype=torch.float64, device="cuda", generator=gen) - 0.5
    W0 = torch.rand(o, o, o, o, dtype=torch.float64, device="cuda", generator=gen) - 0.5
    e0, e1 = slab_range(v, world, rank)
    m0, m1 = slab_range(o, world, rank)
    g_r, t_r = g[:, :, e0:e1, :].contiguous(), t[:, :, e0:e1, :].contiguous()
    for mode in ("push", "p2p", "nccl"):
        ks = KSplitContraction("mnij", (o, o, o, o), "mnef", "ijef", torch.float64, mode=mode)
        Wl = W0[m0:m1].clone()
        ref = W0[m0:m1].clone()
        for alpha, beta in ((0.25, 1.0), (1.0, 0.0), (-0.5, 2.0)):
            ops.contract("mnij", ref, "mnef", g[m0:m1], "ijef", t, alpha, beta)     # full K on this GPU
            ks(Wl, g_r, t_r, alpha, beta)
            torch.cuda.synchronize()
            err = ((Wl - ref).abs().max() / ref.abs().max()).item()
            assert err <= 1e-12, (mode, alpha, beta, err)
        ks.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_ksplit_modes_two_gpus():
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port()), nprocs=2, join=True)
