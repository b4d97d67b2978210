#!/usr/bin/env python
"""Config 5 of BASELINE.json: Wabef[a,b,e,f] = W + 0.25 * sum_mn g[m,n,e,f] t[m,n,a,b] with o=40, v=400 (W = 204.8 GB),
sharded over the uncontracted index `a` across the ranks of one node (STRONG scaling: total work fixed at 8.192e13 flop).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/run_sharded_wabef.py

Each rank owns a contiguous a-slab (einsums_b200.sharding.slab_range), keeps g and t replicated (2 x 2.05 GB) and walks its
slab in sub-slabs that fit a memory budget, reusing one C buffer (W never exists whole on one device; at N=1 the 400 rows are
produced 100 at a time). No collective in the timed region; one all-reduce(MAX) of the per-rank time afterwards.
A parity spot check compares one row of one sub-slab with a float64 torch.einsum on the same device.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from einsums_b200 import ops
from einsums_b200.sharding import plan_slabs

ap = argparse.ArgumentParser()
ap.add_argument("--o", type=int, default=40)
ap.add_argument("--v", type=int, default=400)
ap.add_argument("--budget-gb", type=float, default=60.0)
ap.add_argument("--reps", type=int, default=1)
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
ops.init(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

o, v = args.o, args.v
plan = plan_slabs(v, v**3, o * o, world, rank, int(args.budget_gb * 2**30))
gen = torch.Generator(device="cuda").manual_seed(20251017 + 5)  # same operands on every rank (replicated)
g = torch.rand(o, o, v, v, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
t = torch.rand(o, o, v, v, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
rows = max(e - s for s, e in plan.sub)
W = torch.zeros(rows, v, v, v, dtype=torch.float64, device="cuda")


def run():
    for s, e in plan.sub:
        ops.contract("abef", W[: e - s], "mnef", g, "mnab", t[:, :, s:e, :], 0.25, 1.0)


run()  # warm-up
torch.cuda.synchronize()
# parity spot check: the first row of this rank's slab, computed alone into a fresh buffer, against a float64
# torch.einsum of the same row on the same device (different summation order) -- tolerance 1e-12 relative
s0 = plan.a0
Wrow = torch.zeros(1, v, v, v, dtype=torch.float64, device="cuda")
ops.contract("abef", Wrow, "mnef", g, "mnab", t[:, :, s0 : s0 + 1, :], 0.25, 1.0)
ref = 0.25 * torch.einsum("mnef,mnb->bef", g, t[:, :, s0, :])
err = ((Wrow[0] - ref).abs().max() / ref.abs().max()).item()
assert err < 1e-12, f"parity failure on rank {rank}: {err}"
del Wrow, ref
W.zero_()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.reps):
    run()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / args.reps], dtype=torch.float64, device="cuda")
mine = ms.item()
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
total_flops = 2.0 * v**4 * o**2
if rank == 0:
    print(json.dumps({"config": f"Wabef o={o} v={v} sharded over a", "n_gpus": world, "ms": ms.item(), "tflops_total": total_flops / ms.item() / 1e9,
                      "tflops_per_gpu": total_flops / ms.item() / 1e9 / world, "rank0_ms": mine, "rows_rank0": plan.a1 - plan.a0,
                      "sub_slabs_rank0": len(plan.sub), "parity_rel_err_row": err, "collectives_in_timed_region": 0}), flush=True)
if world > 1:
    dist.destroy_process_group()
