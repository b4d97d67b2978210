#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/probe_z_packed.py 2>&1 | tee gpurun_out/r2y_z_packed.jsonl
timeout 300 python tools/probe_c64.py 2>&1 | tee gpurun_out/r2y_c64.jsonl
timeout 600 python -m pytest tests -m gpu -q -x -k "c128 or zgemm or fuzz or small_all or conj or tcgen05" > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2y_pytest.log | cut -c1-600
