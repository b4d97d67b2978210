#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "tcgen05" > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2y_pytest.log | cut -c1-600
(for m in 1 0; do echo "pair=$m"; timeout 200 python tools/probe_f32_generic.py $m; timeout 200 python tools/probe_c64.py $m; done) 2>&1 | grep -v "^$" | cut -c1-200 | tee gpurun_out/r2y_generic_pair.txt
