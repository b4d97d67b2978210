#!/usr/bin/env python
"""K-split sharding (a CONTRACTED index distributed over the GPUs) on the CCD Wmnij pattern at vvvv-sized operands:
    Wmnij[m,n,i,j] = W + 0.25 * sum_ef g[m,n,e,f] t[i,j,e,f],  o=40, v=400: operands 2 x 2.05 GB, C = 20.5 MB, 8.19e11 flop
`e` is split across the ranks; C stays sharded along m. Compares the fused peer-memory path (GEMM epilogue stores straight into the owner's
receive slot over NVLink, then a fixed-order slot sum) with the NCCL fallback (reduce to the owner) and with the compute alone.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29551 tools/run_ksplit.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from einsums_b200 import ops
from einsums_b200.sharding import KSplitContraction, slab_range

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
ops.init(local)
os.dup2(2, 1) if False else None
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
o, v = int(os.environ.get("KS_O", 40)), int(os.environ.get("KS_V", 400))
gen = torch.Generator(device="cuda").manual_seed(20251017 + 6)
g = torch.rand(o, o, v, v, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
t = torch.rand(o, o, v, v, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
W0 = torch.rand(o, o, o, o, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
e0, e1 = slab_range(v, world, rank)
m0, m1 = slab_range(o, world, rank)
g_r, t_r = g[:, :, e0:e1, :].contiguous(), t[:, :, e0:e1, :].contiguous()
flops = 2.0 * o**4 * v**2

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize(); dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(reps): fn()
    ev1.record(); torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1) / reps], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item()

# reference rows of this rank from the FULL operands on this GPU (what a single GPU would compute)
ref = W0[m0:m1].clone()
ops.contract("mnij", ref, "mnef", g[m0:m1], "ijef", t, 0.25, 1.0)
ms_single = timed(lambda: ops.contract("mnij", torch.empty_like(W0), "mnef", g, "ijef", t, 0.25, 0.0), 3)
out = {"ksplit": True, "workload": f"Wmnij o={o} v={v} FP64, e split over {world} ranks, C sharded along m", "n_gpus": world, "flops": flops,
       "single_gpu_full_k_ms": ms_single, "single_gpu_tflops": flops / ms_single / 1e9}
for mode in ("p2p", "p2p_grouped", "push", "nccl"):
    try:
        ks = KSplitContraction("mnij", (o, o, o, o), "mnef", "ijef", torch.float64, mode=mode.split("_")[0], grouped=(mode == "p2p_grouped"))
    except Exception as e:  # noqa: BLE001
        out[mode] = {"error": f"{type(e).__name__}: {e}"[:300]}
        continue
    Wl = W0[m0:m1].clone()
    ks(Wl, g_r, t_r, 0.25, 1.0)
    torch.cuda.synchronize()
    err = torch.tensor([((Wl - ref).abs().max() / ref.abs().max()).item()], dtype=torch.float64, device="cuda")
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    Wl2 = W0[m0:m1].clone()
    ks(Wl2, g_r, t_r, 0.25, 1.0)
    same = torch.tensor([float(torch.equal(Wl, Wl2))], device="cuda"); dist.all_reduce(same, op=dist.ReduceOp.MIN)
    ms = timed(lambda: ks(Wl, g_r, t_r, 0.25, 0.0))
    ks.profile = []
    dist.barrier(); ks(Wl, g_r, t_r, 0.25, 0.0); torch.cuda.synchronize()
    ph = {ks.profile[i + 1][0]: round(ks.profile[i][1].elapsed_time(ks.profile[i + 1][1]), 3) for i in range(len(ks.profile) - 1)}
    ks.profile = None
    allph = [None] * world
    dist.all_gather_object(allph, ph)
    out[mode] = {"ms": ms, "tflops_total": flops / ms / 1e9, "speedup_vs_single": ms_single / ms, "parity_rel_err": err.item(), "bit_reproducible": bool(same.item()),
                 "kernel": ops.last_kernel(), "phases_ms_per_rank": allph}
    ks.close()
# compute alone: the same per-destination partial products into local memory, no exchange
scratch = torch.empty(o, o, o, o, dtype=torch.float64, device="cuda")
def compute_only():
    Cs, As, Bs = [], [], []
    for d in range(world):
        lo, hi = slab_range(o, world, d)
        if hi > lo: Cs.append(scratch[lo:hi]); As.append(g_r[lo:hi]); Bs.append(t_r)
    ops.contract_grouped("mnij", Cs, "mnef", As, "ijef", Bs, 0.25, 0.0)
ms_c = timed(compute_only)
out["compute_only_ms"] = ms_c
ms_one = timed(lambda: ops.contract("mnij", scratch, "mnef", g_r, "ijef", t_r, 0.25, 0.0))
out["compute_one_launch_local_ms"] = ms_one   # the rank's whole partial product as ONE contraction into local memory (no per-destination split)
if rank == 0:
    print(json.dumps(out), flush=True)
dist.destroy_process_group()
