#!/bin/bash
# round 2, first GPU call: smoke, parity suite, host topology, bench, A/B of the multi-index TMA kernel
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2a_pytest.log
timeout 300 python tools/host_copy_ceiling.py > gpurun_out/r2a_host.json 2> gpurun_out/r2a_host.err; echo "host rc=$?"
timeout 900 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2a_bench.err
EB200_USE_TMA_GENERIC=0 timeout 600 python tools/run_configs.py 2 3 > gpurun_out/r2a_cfg_gather.jsonl 2>&1; echo "cfg gather rc=$?"
EB200_TMA_TILE=0 timeout 300 python tools/run_configs.py 2 > gpurun_out/r2a_cfg_tile0.jsonl 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt
