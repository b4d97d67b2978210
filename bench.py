#!/usr/bin/env python
"""bench.py — headline benchmark of the contraction hot path (contract: see the task brief §④).

Workload (BASELINE.json configs[1]): CCD W intermediates, o=20, v=200, FP64, one step =
    Wmnij[m,n,i,j] = 1.0*Wmnij + 0.25 * sum_ef g[m,n,e,f] t[i,j,e,f]      (M=N=400,   K=40000, split-K)
    Wabef[a,b,e,f] = 1.0*Wabef + 0.25 * sum_mn g[m,n,e,f] t[m,n,a,b]      (M=N=40000, K=400, C = 12.8 GB)
both through the C ABI (eb200_contract -> gemm_tma_f64: TMA + mbarrier + DMMA). flops/step = 2*o^4*v^2 + 2*v^4*o^2 = 1.2928e12.

N > 1 (torchrun, one rank per GPU): the `a` index of Wabef is sharded — rank r owns the slab a in [r*v,(r+1)*v) of a
W with N*v values of `a` (per-GPU work fixed => weak scaling): g_oovv is the same on every rank, t[:,:,a_r,:] and the W
slab are the rank's own; no collective on the data path.

Besides the headline the same JSON line carries
  "sharded": BASELINE configs[4] — Wabef o=40 v=400 (W = 204.8 GB, 8.192e13 flop) sharded over `a` across the N ranks with
             both operands replicated (identical seeds): STRONG scaling, total work fixed, at every N including 1;
  "configs": (N = 1 only) the other BASELINE configs at full size — Fock J/K n=100, Wmnij, Wmbej o=60 v=240, Fock J/K and
             permuted-K at n=300, the 24-order rank-4 permute sweep — each with its roofline fraction (tools/run_configs.py).

--impl reference : the reference's CPU path (oracle port: analyser + OpenBLAS dgemm, all host threads) on the same config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

O, V = 20, 200
FLOPS_WMNIJ = 2.0 * O**4 * V**2
FLOPS_WABEF = 2.0 * V**4 * O**2
FLOPS_STEP = FLOPS_WMNIJ + FLOPS_WABEF
NOMINAL_FP64_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12  # 37.2: 64 FP64 FMA/clk/SM at clocks.max.sm


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-sample-v", type=int, default=0, help="v for the CPU sample (0 = full config)")
    ap.add_argument("--e2e-slab-mb", type=int, default=128, help="slab size of the host-path pipeline (MiB)")
    ap.add_argument("--no-sharded", action="store_true", help="skip the strong-scaling record (Wabef o=40 v=400 sharded over a)")
    ap.add_argument("--no-configs", action="store_true", help="skip the records of the other BASELINE configs (N = 1 only)")
    ap.add_argument("--sharded-budget-gb", type=float, default=60.0, help="largest C sub-slab kept on one GPU")
    ap.add_argument("--cpu-configs", action="store_true",
                    help="with --impl reference: also time the CPU port on the other BASELINE configs at host-sized samples (one JSON line each)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "power_w_max": max(pw) if pw else None, "samples": len(sm)}


def fp64_peak():
    """FP64 tensor-pipe (DMMA) peak measured on this GPU by tools/fp64_peaks; MEASURED_PEAKS.json holds no FP64 figure."""
    exe = os.path.join(ROOT, "tools", "fp64_peaks")
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120, env=dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("LOCAL_RANK", "0")
                                                                                       if "CUDA_VISIBLE_DEVICES" not in os.environ else os.environ["CUDA_VISIBLE_DEVICES"]))
        j = json.loads(out.stdout.strip().splitlines()[-1])
        return float(j["dmma_tflops"]), "measured in this run by tools/fp64_peaks (DMMA.8x8x4 issue-bound loop)", j
    except Exception as e:  # noqa: BLE001
        return NOMINAL_FP64_TFLOPS, f"fallback: nominal 148 SM x 64 FMA/clk x 1.965 GHz ({type(e).__name__})", None


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference(steps, warmup, v_sample=0, threads=None):
    import numpy as np
    from oracle import binding as orc
    info = orc.load_openblas()
    ncores = os.cpu_count() or 1
    threads = threads or ncores
    orc.lib().orc_blas_set_threads(threads)
    o, v = O, (v_sample or V)
    rng = np.random.default_rng(20251017 + 2)
    g = rng.uniform(-1, 1, (o, o, v, v))
    t = rng.uniform(-1, 1, (o, o, v, v))
    Wm = rng.uniform(-1, 1, (o, o, o, o))
    Wa = rng.uniform(-1, 1, (v, v, v, v))
    flops = 2.0 * o**4 * v**2 + 2.0 * v**4 * o**2

    def step():
        a1 = orc.einsum_f64("mnij", Wm, "mnef", g, "ijef", t, 0.25, 1.0)
        a2 = orc.einsum_f64("abef", Wa, "mnef", g, "mnab", t, 0.25, 1.0)
        assert a1 == "GEMM" and a2 == "GEMM", (a1, a2)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return {"value": flops / dt / 1e12, "unit": "TFLOP/s", "cores": threads, "kind": "port",
            "sample": (f"full config o={o} v={v}" if v == V else f"reduced sample o={o} v={v} (same two einsums)") +
                      f", {steps} steps after {warmup} warm-up; oracle analyser -> {info}", "ms_per_step": dt * 1e3}


def cpu_other_configs():
    """SURVEY.md section 8(d) 'CPU baseline beside it' for the non-headline configs, at sizes that fit the host: the oracle's
    analyser + OpenBLAS routes / OpenMP generic loops / blocked transpose, all host threads, median of 3 after one warm-up."""
    import numpy as np
    from oracle import binding as orc
    info = orc.load_openblas()
    ncores = os.cpu_count() or 1
    orc.lib().orc_blas_set_threads(ncores)
    rng = np.random.default_rng(20251017)

    def med(fn, reps=3):
        fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return sorted(ts)[len(ts) // 2]

    def emit(cfg, what, value, unit, route, sample):
        print(json.dumps({"impl": "reference", "config": {"workload": cfg}, "op": what, "value": value, "unit": unit, "route": route,
                          "cpu_baseline": {"value": value, "unit": unit, "cores": ncores, "kind": "port", "sample": sample + f"; {info}"}}), flush=True)

    n = 100
    g = rng.uniform(-1, 1, (n, n, n, n)); D = rng.uniform(-1, 1, (n, n)); J = np.zeros((n, n)); K = np.zeros((n, n))
    route = [None]
    t = med(lambda: route.__setitem__(0, orc.einsum_f64("pq", J, "pqrs", g, "rs", D, 2.0, 0.0)))
    emit("cfg1 Fock J n=100", "pq<-pqrs,rs", 8.0 * n**4 / t / 1e9, "GB/s", route[0], "full size")
    t = med(lambda: route.__setitem__(0, orc.einsum_f64("pq", K, "prqs", g, "rs", D, -1.0, 0.0)))
    emit("cfg1/cfg3b Fock K n=100", "pq<-prqs,rs", 8.0 * n**4 / t / 1e9, "GB/s", route[0], "full size for cfg1; cfg3b (n=300, 64.8 GB) not sampled")
    del g
    o, v = 30, 120
    t2 = rng.uniform(-1, 1, (o, o, v, v)); gg = rng.uniform(-1, 1, (o, o, v, v)); W = rng.uniform(-1, 1, (o, v, v, o))
    t = med(lambda: route.__setitem__(0, orc.einsum_f64("mbej", W, "jnfb", t2, "mnef", gg, -0.5, 1.0)), reps=1)
    emit("cfg3a Wmbej", "mbej<-jnfb,mnef", 2.0 * o**3 * v**3 / t / 1e12, "TFLOP/s", route[0], f"reduced sample o={o} v={v} (config: o=60 v=240), 1 rep")
    for n in (64, 96):
        A = rng.uniform(-1, 1, (n, n, n, n)); Cn = np.empty_like(A)
        for order in ("lkji", "ikjl", "jilk"):
            t = med(lambda: orc.permute(order, Cn, "ijkl", A, 1.0, 0.0, fast=True))
            emit(f"cfg4 permute n={n}", f"{order}<-ijkl", 16.0 * n**4 / t / 1e9, "GB/s", "blocked transpose", "full size, 3 of the 24 orders")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    warm = 1 if args.warmup > 0 else 0
    cb = cpu_reference(steps, warm, args.cpu_sample_v)
    line = {
        "impl": "reference", "metric": "FP64 einsum TFLOP/s", "value": cb["value"], "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic uniform[-1,1), seeded", "config": {"workload": f"CCD Wmnij+Wabef o={O} v={V} FP64 (BASELINE configs[1])", "o": O, "v": V},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": f"CPU arm: one step takes about a second on the host, so it runs at most 3 timed steps after 1 warm-up whatever --steps/--warmup "
                f"say (asked: {args.steps}/{args.warmup}; ran: {steps}/{warm}); value is per step, independent of the count",
    }
    print(json.dumps(line), flush=True)
    if args.cpu_configs:
        cpu_other_configs()


# ------------------------------------------------------------------------------------------------ our arm
def bind_to_gpu_numa_node(local):
    """Run this rank on the CPUs of the NUMA node its GPU hangs off, so that first-touch pinned allocations land on that node
    (the e2e path is bound by host memory / PCIe root bandwidth, not by a kernel). Returns what was done, for the JSON line."""
    import glob
    import torch
    nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))
    info = {"numa_nodes_visible": len(nodes)}
    try:
        pr = torch.cuda.get_device_properties(local)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip()
        info["gpu_numa_node"] = node
        if len(nodes) > 1 and node not in ("-1", ""):
            cpus = set()
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                info["bound_to_cpus"] = len(cpus)
    except Exception as e:  # noqa: BLE001
        info["error"] = f"{type(e).__name__}: {e}"[:120]
    return info


def run_sharded(args, world, rank, barrier):
    """BASELINE configs[4]: Wabef[a,b,e,f] = W + 0.25 sum_mn g[m,n,e,f] t[m,n,a,b], o=40, v=400 (W = 204.8 GB, 8.192e13 flop), the
    a-index block-distributed over the ranks (einsums_b200.sharding), g and t REPLICATED (same seed on every rank; 2 x 2.05 GB),
    each rank walking its slab in sub-slabs of at most --sharded-budget-gb through one reused C buffer. Strong scaling: total
    work is fixed. No collective in the timed region; max over ranks of the CUDA-event time."""
    import torch
    import torch.distributed as dist
    from einsums_b200 import ops
    from einsums_b200.sharding import plan_slabs
    o, v = 40, 400
    plan = plan_slabs(v, v**3, o * o, world, rank, int(args.sharded_budget_gb * 2**30))
    gen = torch.Generator(device="cuda").manual_seed(20251017 + 5)   # identical on every rank: the operands are replicated
    g = torch.rand(o, o, v, v, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
    t = torch.rand(o, o, v, v, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
    check = torch.stack([g.sum(), t.sum()])
    if world > 1:   # prove the replication: the operand checksums of all ranks must coincide
        lo, hi = check.clone(), check.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "g/t differ between ranks"
    rows = max(e - s for s, e in plan.sub)
    W = torch.zeros(rows, v, v, v, dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream()

    def run():
        for s, e in plan.sub:
            ops.contract("abef", W[: e - s], "mnef", g, "mnab", t[:, :, s:e, :], 0.25, 1.0)

    run()   # warm-up (also the first-touch of W)
    # parity: the first row of this rank's slab against a float64 torch.einsum on the same device (different summation order)
    s0 = plan.a0
    Wrow = torch.zeros(1, v, v, v, dtype=torch.float64, device="cuda")
    ops.contract("abef", Wrow, "mnef", g, "mnab", t[:, :, s0:s0 + 1, :], 0.25, 1.0)
    ref = 0.25 * torch.einsum("mnef,mnb->bef", g, t[:, :, s0, :])
    err = torch.tensor([((Wrow[0] - ref).abs().max() / ref.abs().max()).item()], dtype=torch.float64, device="cuda")
    del Wrow, ref
    W.zero_()
    l0 = ops.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    run()
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
    flops = 2.0 * v**4 * o**2
    out = {"workload": f"Wabef o={o} v={v} FP64 a-sharded (BASELINE configs[4]); W = {8 * v**4 / 1e9:.1f} GB produced in sub-slabs of <= {args.sharded_budget_gb:g} GB",
           "scaling": "strong", "n_gpus": world, "ms": ms.item(), "tflops_total": flops / ms.item() / 1e9, "tflops_per_gpu": flops / ms.item() / 1e9 / world,
           "rows_per_rank": plan.a1 - plan.a0, "sub_slabs_per_rank": len(plan.sub), "parity_rel_err": err.item(), "parity_tol": 1e-12,
           "operands": "g, t replicated: same generator seed on every rank, checksums compared across ranks", "collectives_in_timed_region": 0,
           "gpu_launches": int(ops.launch_count() - l0), "kernel": ops.last_kernel()}
    assert out["parity_rel_err"] < 1e-12, out
    del W
    torch.cuda.empty_cache()
    if world > 1:
        out["ksplit"] = run_ksplit(world, rank, barrier, g, t, o, v)
    del g, t
    torch.cuda.empty_cache()
    return out


def run_ksplit(world, rank, barrier, g, t, o, v):
    """SURVEY 8(e) row 2 — a CONTRACTED index distributed: Wmnij[m,n,i,j] = 0.25 sum_ef g[m,n,e,f] t[i,j,e,f] at o=40 v=400 (8.19e11 flop,
    operands 2 x 2.05 GB, C = 20.5 MB) with `e` split over the ranks and C sharded along m (einsums_b200.sharding.KSplitContraction,
    mode "push": one local product, peer-to-peer copies of the row slabs into the owners' receive slots over NVLink, stream-ordered
    barrier, fixed-order slot sum). Timed region = product + exchange + sum; max over ranks; parity against the full-K product of this
    rank's rows on one GPU."""
    import torch
    import torch.distributed as dist
    from einsums_b200 import ops
    from einsums_b200.sharding import KSplitContraction, slab_range
    e0, e1 = slab_range(v, world, rank)
    m0, m1 = slab_range(o, world, rank)
    g_r, t_r = g[:, :, e0:e1, :].contiguous(), t[:, :, e0:e1, :].contiguous()
    ref = torch.zeros(m1 - m0, o, o, o, dtype=torch.float64, device="cuda")
    ops.contract("mnij", ref, "mnef", g[m0:m1], "ijef", t, 0.25, 0.0)
    flops = 2.0 * o**4 * v**2
    stream = torch.cuda.current_stream()

    def timed(fn, reps):
        fn(); barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
        barrier()
        ms = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    full = torch.empty(o, o, o, o, dtype=torch.float64, device="cuda")
    ms_single = timed(lambda: ops.contract("mnij", full, "mnef", g, "ijef", t, 0.25, 0.0), 3)
    del full
    try:
        ks = KSplitContraction("mnij", (o, o, o, o), "mnef", "ijef", torch.float64, mode="push")
        mode = "push (peer-to-peer copies into the owners' slots, CUDA IPC)"
    except Exception as e:  # noqa: BLE001  -- no peer access between the processes: NCCL reduce to the owner
        ks = KSplitContraction("mnij", (o, o, o, o), "mnef", "ijef", torch.float64, mode="nccl")
        mode = f"nccl (peer memory unavailable: {type(e).__name__})"
    Wl = torch.zeros(m1 - m0, o, o, o, dtype=torch.float64, device="cuda")
    l0 = ops.launch_count()
    ks(Wl, g_r, t_r, 0.25, 0.0)
    launches = int(ops.launch_count() - l0)
    torch.cuda.synchronize()
    err = torch.tensor([((Wl - ref).abs().max() / ref.abs().max()).item()], dtype=torch.float64, device="cuda")
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    ms = timed(lambda: ks(Wl, g_r, t_r, 0.25, 0.0), 5)
    ks.close()
    out = {"workload": f"Wmnij o={o} v={v} FP64, contracted index e split over {world} ranks, C sharded along m", "mode": mode, "ms": ms,
           "tflops_total": flops / ms / 1e9, "single_gpu_full_k_ms": ms_single, "speedup_vs_single_gpu": ms_single / ms, "parity_rel_err": err.item(),
           "parity_tol": 1e-12, "gpu_launches_per_call": launches, "collectives_in_timed_region": "two one-element NCCL all-reduces per call (stream-ordered barriers)"}
    assert out["parity_rel_err"] < 1e-12, out
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from einsums_b200 import KF, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: the backend has no CPU fallback"
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local)
    ops.init(local)
    if world > 1:
        # stdout carries exactly one JSON line: while the communicator comes up, NCCL's banner (NCCL_DEBUG=VERSION prints
        # "NCCL version ..." with printf) is sent to stderr by pointing fd 1 at fd 2
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    o, v = O, V
    gen_all = torch.Generator(device="cuda").manual_seed(20251017 + 2)          # same stream on every rank: replicated tensors
    gen = torch.Generator(device="cuda").manual_seed(20251017 + 2 + 1000 * (rank + 1))  # this rank's own slabs
    rnd = lambda *s, gen=gen: torch.rand(*s, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
    g = rnd(o, o, v, v, gen=gen_all)               # g_oovv: replicated (identical on every rank)
    t = rnd(o, o, v, v)                            # t[m,n,a_r,b]: this rank's a-slab of the N*v-wide t
    Wm, Wa = rnd(o, o, o, o), rnd(v, v, v, v)      # Wmnij (per rank) and the Wabef slab a in [rank*v, (rank+1)*v)
    stream = torch.cuda.current_stream()

    kinds, knames = {}, {}

    def step(ev=None):
        d1 = ops.contract("mnij", Wm, "mnef", g, "ijef", t, 0.25, 1.0)
        knames["wmnij"] = ops.last_kernel()
        if ev is not None:
            ev[0].record(stream)
        d2 = ops.contract("abef", Wa, "mnef", g, "mnab", t, 0.25, 1.0)
        knames["wabef"] = ops.last_kernel()
        if ev is not None:
            ev[1].record(stream)
        kinds["wmnij"], kinds["wabef"] = d1, d2

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    K = args.steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    mids = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    ops.launch_count_reset()
    barrier()
    e0.record(stream)
    for i in range(K):
        step(mids[i])
    e1.record(stream)
    barrier()
    launches = ops.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    ms_wabef = sum(m[0].elapsed_time(m[1]) for m in mids) / K
    tms = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_step = tms.item() / K
    value = world * FLOPS_STEP / (ms_step * 1e-3) / 1e12

    # ---- end to end with HOST buffers (pinned): every step uploads all inputs and brings both results back.
    # Wabef (12.8 GB in, 12.8 GB out) goes through the C ABI's eb200_contract_host: slabs along `a`,
    # H2D | contract | D2H overlapped on three streams. Wmnij (1.3 MB) is copied whole.
    e2e = None
    if not args.no_e2e:
        def pinned_copy(x):
            h = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
            h.copy_(x)
            return h
        hg, ht, hWm = pinned_copy(g), pinned_copy(t), pinned_copy(Wm)
        oWm = torch.empty(hWm.shape, dtype=hWm.dtype, pin_memory=True)
        # Wabef lives in ONE pinned buffer per rank (12.8 GB; 8 ranks share the node's RAM): the result overwrites the
        # input slab by slab, so successive steps keep accumulating — same bytes moved, same flops.
        hWa = pinned_copy(Wa)
        oWa = hWa
        ref_rows = hWa[:2].clone()
        del Wa
        torch.cuda.empty_cache()
        counts = {}
        ge, te_ = torch.empty_like(g), torch.empty_like(t)       # device copies refreshed from the host every step

        def e2e_step():
            ge.copy_(hg, non_blocking=True)
            te_.copy_(ht, non_blocking=True)
            Wm.copy_(hWm, non_blocking=True)
            ops.contract("mnij", Wm, "mnef", ge, "ijef", te_, 0.25, 1.0)
            oWm.copy_(Wm, non_blocking=True)
            # the C-ABI host path: Wabef stays in host memory, slabs in and out overlapped with the contraction
            st = ops.contract_host("abef", hWa, "mnef", ge, "mnab", te_, 0.25, 1.0, slab_bytes=args.e2e_slab_mb << 20)
            counts["h2d"] = st.h2d_bytes + (hg.numel() + ht.numel() + hWm.numel()) * 8
            counts["d2h"] = st.d2h_bytes + oWm.numel() * 8

        e2e_step(); barrier()
        ke = max(1, min(K, 3))
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(ke):
            e2e_step()
        a1.record(stream)
        barrier()
        te = torch.tensor([a0.elapsed_time(a1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        # parity of the pipelined path against the resident path on a sample of rows (same inputs, beta = 1)
        chk = ref_rows.cuda()
        for _ in range(ke + 1):   # the in-place host buffer received ke + 1 accumulations
            ops.contract("abef", chk, "mnef", g, "mnab", t[:, :, :2, :], 0.25, 1.0)
        e2e_err = (chk.cpu() - oWa[:2]).abs().max().item()
        e2e = {"value": world * FLOPS_STEP / (te.item() / ke * 1e-3) / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": counts["h2d"],
               "d2h_bytes_per_step": counts["d2h"], "ms_per_step": te.item() / ke, "steps": ke, "max_abs_diff_vs_resident_path": e2e_err,
               "note": "pinned host -> device of g, t, Wmnij, Wabef; both einsums; device -> pinned host of both results; "
                       f"Wabef through eb200_contract_host in {args.e2e_slab_mb} MiB slabs with H2D|compute|D2H overlapped"}
        del e2e_step, hWa, oWa, hg, ht, hWm, oWm, ge, te_, chk   # 12.8 GB of pinned host memory per rank goes back before the next section
        # What the host allows: a bare pinned-memory copy loop, H2D and D2H at the same time on two streams, ALL ranks at once
        # (1 GiB each way per rank, 6 repetitions). e2e moves h2d + d2h bytes per step per rank; its rate is to be read against this.
        n8 = (1 << 30) // 8
        hb_in = torch.empty(n8, dtype=torch.float64, pin_memory=True).fill_(1.0)
        hb_out = torch.empty(n8, dtype=torch.float64, pin_memory=True)
        db_in, db_out = torch.empty(n8, dtype=torch.float64, device="cuda"), torch.ones(n8, dtype=torch.float64, device="cuda")
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

        def both():
            with torch.cuda.stream(s1):
                db_in.copy_(hb_in, non_blocking=True)
            with torch.cuda.stream(s2):
                hb_out.copy_(db_out, non_blocking=True)

        both(); barrier()
        t0 = time.perf_counter()
        for _ in range(6):
            both()
        barrier()
        dtc = time.perf_counter() - t0
        rate = torch.tensor([6 * 2 * n8 * 8 / dtc / 1e9], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(rate, op=dist.ReduceOp.SUM)
        e2e_rate = world * (counts["h2d"] + counts["d2h"]) / (e2e["ms_per_step"] * 1e-3) / 1e9
        e2e["numa"] = numa
        e2e["host_copy_ceiling"] = {"aggregate_both_directions_gbs": rate.item(), "e2e_copy_rate_gbs": e2e_rate, "frac": e2e_rate / rate.item(),
                                    "how": "bare pinned H2D + D2H copy loop, 1 GiB each way per rank, all ranks at once, wall clock between barriers"}
        del hb_in, hb_out, db_in, db_out

    # ---- BASELINE configs[4]: Wabef o=40 v=400 sharded over `a` (strong scaling), at every N
    del g, t, Wm
    Wa = None
    torch.cuda.empty_cache()
    sharded = None
    if not args.no_sharded:
        sharded = run_sharded(args, world, rank, barrier)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src, peak_json = fp64_peak()
    if sharded:
        sharded["frac_of_dmma"] = sharded["tflops_total"] / world / peak
    achieved = FLOPS_WABEF / (ms_wabef * 1e-3) / 1e12
    traffic = None
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get("gemm_tma_f64_wabef_dram_bytes")
        except Exception:
            traffic = None
    line = {
        "metric": "FP64 einsum TFLOP/s", "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic uniform[-1,1), seeded on device",
        "config": {"workload": f"CCD Wmnij+Wabef o={O} v={V} FP64 (BASELINE configs[1])", "o": O, "v": V,
                   "flops_per_step_per_gpu": FLOPS_STEP, "l2": "per-step traffic (25.9 GB) exceeds the 126 MB L2; no flush needed",
                   "sharding": "none" if world == 1 else f"a-index of Wabef sharded over {world} ranks, no collective"},
        "roofline": {"bound": "tensor", "kernel": "gemm_tma_f64 (Wabef, M=N=40000 K=400)", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "ms_per_launch": ms_wabef,
                     "algorithmic_flops_per_launch": FLOPS_WABEF},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "kernels": {k: {"family": int(d.kernel_used), "split_k": int(d.split_k), "launches": int(d.launches)} for k, d in kinds.items()},
        "kernel_names": knames,
    }
    if peak_json:
        line["fp64_peaks"] = peak_json
    if sharded:
        line["sharded"] = sharded
    if not args.no_configs and world == 1:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import run_configs
        torch.cuda.empty_cache()
        csampler = ClockSampler(local)
        csampler.start()
        l0 = ops.launch_count()
        records, pk = run_configs.run_all(("1", "2", "3", "4", "T"), wabef=False)   # Wabef is the headline above
        line["configs"] = {"records": records, "peaks": pk, "clocks": csampler.stop(), "gpu_launches": int(ops.launch_count() - l0),
                           "timing": "median of CUDA-event timings after warm-up; ops under 1 ms are timed as 20 back-to-back launches; inputs exceed L2 "
                                     "(except cfg1's D and Wmnij's 1.3 MB output)"}
    if not args.no_cpu_baseline and world == 1:
        torch.cuda.empty_cache()
        cb = cpu_reference(2, 1, args.cpu_sample_v)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
