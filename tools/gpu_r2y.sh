#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "packed or tma_generic or fuzz" 2>&1 | tail -4 | cut -c1-400
python tools/probe_wmbej.py 61 243 1
