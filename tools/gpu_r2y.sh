#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "pair_stress" 2>&1 | tail -12 | cut -c1-500
