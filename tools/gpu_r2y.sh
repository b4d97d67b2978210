#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "tcgen05 or sgemm" 2>&1 | tail -3 | cut -c1-400
timeout 200 python tools/probe_f32_gap.py 2>&1 | head -3
timeout 120 python tools/probe_f32.py 1 2>&1 | tail -2
