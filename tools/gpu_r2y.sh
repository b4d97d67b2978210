#!/bin/bash
mkdir -p gpurun_out
K="tcgen05_generic_patterns or c64_tcgen05_patterns or c128_packed or f32_packed or tcgen05_layouts or tcgen05_persistent"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "$K" > gpurun_out/r2y_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r2y_memcheck.log | cut -c1-300
