#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2k_pytest.log | cut -c1-300
./build_cpp/test_device > gpurun_out/r2k_cpp.log 2>&1; tail -30 gpurun_out/r2k_cpp.log
./build_cpp/bench_tiled > gpurun_out/r2k_bench_tiled.jsonl 2>&1; cat gpurun_out/r2k_bench_tiled.jsonl
