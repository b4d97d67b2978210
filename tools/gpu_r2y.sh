#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "shim" 2>&1 | tail -15 | cut -c1-600
