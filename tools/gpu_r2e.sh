#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "tcgen05 or sgemm" > gpurun_out/r2e_pytest_tf32.log 2>&1; echo "pytest tf32 rc=$?"; tail -15 gpurun_out/r2e_pytest_tf32.log
timeout 300 python tools/probe_f32.py > gpurun_out/r2e_f32.txt 2>&1; cat gpurun_out/r2e_f32.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2e_pytest.log
for d in 2 3; do echo "gemm dist=$d"; EB200_TMA_DIST=$d python tools/probe.py wabef gemm; done > gpurun_out/r2e_dist.txt 2>&1; cat gpurun_out/r2e_dist.txt
python tools/probe_permute.py ijkl 160; python tools/probe_permute.py ijkl 96
