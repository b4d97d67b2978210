#!/bin/bash
mkdir -p gpurun_out
python tools/probe_level12.py > gpurun_out/r2m_level12.jsonl 2>&1; cat gpurun_out/r2m_level12.jsonl
timeout 600 python -m pytest tests -m gpu -q -x -k "level1 or ger or outer or direct or elementwise or fuzz" > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2m_pytest.log | cut -c1-200
