#!/bin/bash
mkdir -p gpurun_out
timeout 300 bash tools/ncu_capture.sh r2_dger elementwise 2 python tools/probe_ger.py
