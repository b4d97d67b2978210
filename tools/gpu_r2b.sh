#!/bin/bash
# round 2, second GPU call: full parity suite, ncu of the multi-index TMA kernel on Wmbej and of the worst permute orders
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2b_pytest.log
python tools/probe_wmbej.py 40 160 > gpurun_out/r2b_wmbej_time.txt 2>&1; cat gpurun_out/r2b_wmbej_time.txt
timeout 600 bash tools/ncu_capture.sh r2_wmbej_tma contract_tma_f64 1 python tools/probe_wmbej.py 40 160
for o in lkji ijlk ilkj; do python tools/probe_permute.py $o 160; python tools/probe_permute.py $o 96; done > gpurun_out/r2b_permute_time.txt 2>&1; cat gpurun_out/r2b_permute_time.txt
timeout 300 bash tools/ncu_capture.sh r2_permute_lkji_160 permute 2 python tools/probe_permute.py lkji 160
timeout 300 bash tools/ncu_capture.sh r2_permute_ilkj_160 permute 2 python tools/probe_permute.py ilkj 160
timeout 300 bash tools/ncu_capture.sh r2_permute_ikjl_160 permute 2 python tools/probe_permute.py ikjl 160
