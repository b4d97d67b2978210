#!/usr/bin/env python
"""Summarise an exported `ncu --page raw --csv` file: the metrics the roofline discussion needs."""
import csv, sys, re
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
    "sm__inst_executed_pipe_fp64", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed_pipe_fp64",
    "sm__inst_executed_pipe_tensor", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warp_latency_issue_stalled", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
]
def main(path, extra=None):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]; units = rows[1]; vals = rows[2] if len(rows) > 2 else []
    name_i = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
    if name_i is not None and vals: print("kernel:", vals[name_i][:110])
    pat = re.compile(extra) if extra else None
    for i, h in enumerate(hdr):
        if h in KEYS or (pat and pat.search(h)) or "pipe_fp64" in h or "warp_issue_stalled" in h and "pct" in h or "pipe_tensor" in h and "pct" in h:
            if i < len(vals): print(f"  {h} [{units[i]}] = {vals[i]}")
if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
