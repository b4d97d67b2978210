#!/bin/bash
mkdir -p gpurun_out
python tools/probe_tiles.py
python tools/probe_wmbej.py 60 240; python tools/probe_wmbej.py 40 160
python tools/probe.py wmnij
timeout 900 python -m pytest tests -m gpu -q -x -k "tma or contract or fuzz or grouped or tiled or wmnij or wmbej" > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2x_pytest.log | cut -c1-300
