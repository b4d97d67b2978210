#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tools/run_ksplit.py > gpurun_out/r2p_ksplit_n$N.json 2> gpurun_out/r2p_ksplit_n$N.err; echo "rc=$?"
cat gpurun_out/r2p_ksplit_n$N.json | cut -c1-2000; tail -15 gpurun_out/r2p_ksplit_n$N.err | cut -c1-300
