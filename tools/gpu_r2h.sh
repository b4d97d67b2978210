#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -k "zgemm or tcgen05 or sgemm or complex" > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2h_pytest.log | cut -c1-300
timeout 300 python tools/probe_f32.py > gpurun_out/r2h_f32.txt 2>&1; cat gpurun_out/r2h_f32.txt
timeout 300 python tools/probe_z.py > gpurun_out/r2h_z.txt 2>&1; cat gpurun_out/r2h_z.txt
