#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -s -k "default_routing_fuzz" 2>&1 | grep "kernels seen\|passed\|failed"
