#!/bin/bash
mkdir -p gpurun_out
(for m in 1 0; do python tools/probe_wmbej.py 41 163 $m; python tools/probe_wmbej.py 61 243 $m; done; python tools/probe_wmbej.py 60 240 1) 2>&1 | tee gpurun_out/r2y_wmbej_odd.txt
timeout 600 python -m pytest tests -m gpu -q -x -k "tma or fuzz or packed or wmbej or mid" > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2y_pytest.log | cut -c1-600
