#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/probe_f32_generic.py 2>&1 | tee gpurun_out/r2y_f32_generic.jsonl
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2y_pytest.log | cut -c1-600
