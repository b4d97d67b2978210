#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2g_pytest.log | cut -c1-300
timeout 300 python tools/probe_f32.py > gpurun_out/r2g_f32.txt 2>&1; cat gpurun_out/r2g_f32.txt
timeout 600 bash tools/ncu_capture.sh r2_tf32_persistent gemm_tf32x3 1 python tools/probe_f32.py
timeout 600 bash tools/ncu_capture.sh r2_wmbej_tma_dist2 contract_tma_f64 1 python tools/probe_wmbej.py 40 160
timeout 300 bash tools/ncu_capture.sh r2_permute_tma_lkji_160 permute_tma 2 python tools/probe_permute.py lkji 160
timeout 300 bash tools/ncu_capture.sh r2_permute_tma_jkli_160 permute_tma 2 python tools/probe_permute.py jkli 160
