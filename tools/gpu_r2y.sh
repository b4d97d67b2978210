#!/bin/bash
mkdir -p gpurun_out
timeout 300 bash tools/ncu_capture.sh r2_tf32_generic_pair_wmbej gemm_tf32x3 1 python tools/probe_one.py f32wmbej
K="not large and not full_size and not stress and not checksum and not water and not round_trip and not default_routing_fuzz and not persistent and not linearity"
timeout 1100 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "$K" > gpurun_out/r2y_memcheck_all.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2y_memcheck_all.log | cut -c1-300
