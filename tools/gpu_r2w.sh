#!/bin/bash
mkdir -p gpurun_out
timeout 400 bash tools/ncu_capture.sh r2_zgemm gemm_tma_z 1 python tools/probe_z1.py
timeout 300 bash tools/ncu_capture.sh r2_dger_after elementwise_pairs_2d 2 python tools/probe_ger.py
