#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "tcgen05 or sgemm" 2>&1 | tail -4 | cut -c1-400
(echo "pair=1"; timeout 120 python tools/probe_f32.py 1) 2>&1 | tee gpurun_out/r2y_probe_f32_vec4.txt
