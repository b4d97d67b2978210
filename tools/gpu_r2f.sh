#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -k "tcgen05 or sgemm" > gpurun_out/r2f_pytest_tf32.log 2>&1; echo "pytest tf32 rc=$?"; tail -8 gpurun_out/r2f_pytest_tf32.log | cut -c1-300
timeout 300 python tools/probe_f32.py > gpurun_out/r2f_f32.txt 2>&1; cat gpurun_out/r2f_f32.txt
