#!/bin/bash
# compute-sanitizer on the round-2 kernels
mkdir -p gpurun_out
SEL='tma_generic or permute_tma or zgemm_tensor_pipe or zgemm_seam or tcgen05_persistent or tcgen05_layouts'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "$SEL" > gpurun_out/r2o_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/r2o_memcheck.log | tail -3
SEL2='tma_generic_patterns or permute_tma_shapes or zgemm_tensor_pipe or tcgen05_layouts'
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "$SEL2" > gpurun_out/r2o_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/r2o_racecheck.log | tail -3
