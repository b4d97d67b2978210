#!/bin/bash
mkdir -p gpurun_out
python tools/probe_level12.py > gpurun_out/r2u_level12.jsonl 2>&1; cat gpurun_out/r2u_level12.jsonl
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2u_pytest.log | cut -c1-300
./build_cpp/test_device 2>&1 | tail -2
