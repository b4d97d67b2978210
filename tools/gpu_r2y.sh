#!/bin/bash
mkdir -p gpurun_out
timeout 300 bash tools/ncu_capture.sh r2_tf32_wabef_beta1 gemm_tf32x3 1 python tools/probe_wabef_f32_one.py 1.0
timeout 300 bash tools/ncu_capture.sh r2_tf32_wabef_beta0 gemm_tf32x3 1 python tools/probe_wabef_f32_one.py 0.0
