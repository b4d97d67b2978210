#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/probe_f32_layouts.py 2>&1 | tee gpurun_out/r2y_f32_layouts.jsonl
