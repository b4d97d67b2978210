#!/bin/bash
mkdir -p gpurun_out
for g in 8 16 32 64; do echo "group=$g"; EB200_TMA_GROUP=$g python tools/probe.py wabef gemm; done > gpurun_out/r2n_group.txt 2>&1; cat gpurun_out/r2n_group.txt
for g in 8 32; do EB200_TMA_GROUP=$g ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:gemm_tma_f64 -s 2 -c 1 --csv python tools/probe.py wabef 2>/dev/null | grep -E "gemm_tma|Metric" | cut -c1-400 > gpurun_out/r2n_traffic_group$g.csv; cat gpurun_out/r2n_traffic_group$g.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'; done
