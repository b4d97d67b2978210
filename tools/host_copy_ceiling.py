#!/usr/bin/env python
"""Host side of the end-to-end path: where do the GPUs sit (PCIe root / NUMA node), what may this process bind to, and what is
the bare pinned-memory copy ceiling (H2D, D2H, both at once) of ONE rank and of ALL ranks at the same time?

    python tools/host_copy_ceiling.py                              # one GPU: topology + per-NUMA-node copy rates
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/host_copy_ceiling.py

Prints one JSON line ("host_copy_ceiling"); bench.py's multi-rank e2e number is to be read against `aggregate_both_gbs`.
NUMA placement uses the raw set_mempolicy/mbind syscalls (no libnuma in the image); when the container forbids them the
report says so and falls back to the default policy.
"""
import ctypes
import glob
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

libc = ctypes.CDLL(None, use_errno=True)
SYS_mbind, SYS_set_mempolicy, SYS_get_mempolicy = 237, 238, 239  # x86_64
MPOL_DEFAULT, MPOL_PREFERRED, MPOL_BIND = 0, 1, 2


def read(path, default=None):
    try:
        return open(path).read().strip()
    except Exception:
        return default


def set_mempolicy(mode, node):
    mask = ctypes.c_ulong(0 if node is None else (1 << node))
    rc = libc.syscall(SYS_set_mempolicy, ctypes.c_int(mode), ctypes.byref(mask) if node is not None else None, ctypes.c_ulong(64 if node is not None else 0))
    return rc == 0, (os.strerror(ctypes.get_errno()) if rc else "ok")


def topology():
    nodes = {}
    for d in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
        n = int(d.rsplit("node", 1)[1])
        mem = read(d + "/meminfo", "")
        tot = [l for l in mem.splitlines() if "MemTotal" in l]
        free = [l for l in mem.splitlines() if "MemFree" in l]
        nodes[n] = {"cpulist": read(d + "/cpulist"), "mem_total_gb": round(int(tot[0].split()[-2]) / 2**20, 1) if tot else None,
                    "mem_free_gb": round(int(free[0].split()[-2]) / 2**20, 1) if free else None}
    gpus = []
    for dev in sorted(glob.glob("/sys/bus/pci/devices/*")):
        if read(dev + "/vendor") == "0x10de" and (read(dev + "/class", "") or "").startswith("0x0302"):
            gpus.append({"bus": os.path.basename(dev), "numa_node": read(dev + "/numa_node"), "link_speed": read(dev + "/current_link_speed"),
                         "link_width": read(dev + "/current_link_width"), "local_cpulist": read(dev + "/local_cpulist")})
    smi = {}
    for name, cmd in (("topo", ["nvidia-smi", "topo", "-m"]),
                      ("gpus", ["nvidia-smi", "--query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current", "--format=csv,noheader"])):
        try:
            smi[name] = subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout.strip().splitlines()
        except Exception as e:  # noqa: BLE001
            smi[name] = [f"unavailable: {e}"]
    status = read("/proc/self/status", "")
    allowed = {l.split(":")[0]: l.split(":", 1)[1].strip() for l in status.splitlines() if l.startswith(("Cpus_allowed_list", "Mems_allowed_list"))}
    cpu = [l.split(":", 1)[1].strip() for l in (read("/proc/cpuinfo", "") or "").splitlines() if l.startswith("model name")]
    return {"numa_nodes": nodes, "pci_gpus": gpus, "nvidia_smi": smi, "allowed": allowed, "affinity": sorted(os.sched_getaffinity(0))[:4] + ["..."] +
            [len(os.sched_getaffinity(0))], "cpu_model": cpu[0] if cpu else None, "logical_cpus": len(cpu),
            "memtotal_gb": round(int((read("/proc/meminfo", "MemTotal: 0 kB").split()[1])) / 2**20, 1)}


def pinned(nbytes, node):
    """pinned host buffer under mempolicy bind(node) (node None = default policy); returns (tensor, policy status)"""
    ok, why = (True, "default policy") if node is None else set_mempolicy(MPOL_BIND, node)
    t = torch.empty(nbytes // 8, dtype=torch.float64, pin_memory=True)
    t.fill_(1.0)
    if node is not None:
        set_mempolicy(MPOL_DEFAULT, None)
    return t, (why if not ok else "ok")


def rates(h_in, h_out, d_in, d_out, reps=4, barrier=None):
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    out = {}
    for mode in ("h2d", "d2h", "both"):
        if barrier:
            barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        nb = h_in.numel() * 8 * reps * (2 if mode == "both" else 1)
        out[mode] = nb / dt / 1e9
    return out


def main():
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    barrier = (lambda: (dist.barrier(), torch.cuda.synchronize())) if world > 1 else None
    topo = topology() if rank == 0 else None
    nbytes = int(os.environ.get("CEILING_BYTES", str(2 << 30)))
    d_in = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda"); d_out = torch.ones(nbytes // 8, dtype=torch.float64, device="cuda")
    nodes = sorted(int(os.path.basename(d)[4:]) for d in glob.glob("/sys/devices/system/node/node[0-9]*"))
    report = {"host_copy_ceiling": True, "n_gpus": world, "bytes_per_copy": nbytes}
    variants = [("default", None)] + ([(f"bind_node{n}", n) for n in nodes] if len(nodes) > 1 and world == 1 else [])
    if world > 1 and len(nodes) > 1:  # every rank on the node its GPU reports (falls back to round-robin when sysfs says -1)
        props = torch.cuda.get_device_properties(local)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        nn = read(f"/sys/bus/pci/devices/{bus}/numa_node", "-1")
        mine = int(nn) if nn not in (None, "-1") else nodes[local * len(nodes) // max(world, 1)]
        variants.append((f"bind_gpu_node", mine))
    res = {}
    for name, node in variants:
        try:
            h_in, why = pinned(nbytes, node); h_out, _ = pinned(nbytes, node)
        except Exception as e:  # noqa: BLE001
            res[name] = {"error": str(e)[:200]}
            continue
        r = rates(h_in, h_out, d_in, d_out, barrier=barrier)
        t = torch.tensor([r["h2d"], r["d2h"], r["both"]], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t)   # sum over ranks: every rank copies at the same time
        res[name] = {"policy": why, "node": node, "rank0_gbs": {k: round(v, 1) for k, v in r.items()},
                     "aggregate_gbs": {"h2d": round(t[0].item(), 1), "d2h": round(t[1].item(), 1), "both": round(t[2].item(), 1)}}
        del h_in, h_out
        torch._C._host_emptyCache() if hasattr(torch._C, "_host_emptyCache") else None
    report["variants"] = res
    if rank == 0:
        report["topology"] = topo
        print(json.dumps(report), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
