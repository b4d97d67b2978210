#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2j_pytest.log | cut -c1-300
python tools/probe.py wabef gemm > gpurun_out/r2j_probe.txt 2>&1; cat gpurun_out/r2j_probe.txt
timeout 900 python bench.py --no-configs > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"; python -c "
import json; j=json.load(open('gpurun_out/r2j_bench.json')); print(j['value'], j['roofline']['frac'], j['e2e']['value'], j['sharded']['tflops_total'], j['sharded']['frac_of_dmma'])"
