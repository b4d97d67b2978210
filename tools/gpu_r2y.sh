#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/probe_f32_gap.py 2>&1 | tail -3 | tee gpurun_out/r2y_f32_gap2.jsonl
timeout 300 python tools/probe_f32_generic.py 1 2>&1 | cut -c1-220
timeout 300 python tools/probe_c64.py 1 2>&1 | cut -c1-220
timeout 600 python -m pytest tests -m gpu -q -x -k "tcgen05 or packed or fuzz or c64" > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2y_pytest.log | cut -c1-600
