#!/bin/bash
mkdir -p gpurun_out
(echo "pair=1"; timeout 120 python tools/probe_f32.py 1; echo "pair=0"; timeout 120 python tools/probe_f32.py 0) 2>&1 | tee gpurun_out/r2y_probe_f32_pair.txt
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2y_pytest.log | cut -c1-600
