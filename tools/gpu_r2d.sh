#!/bin/bash
# multi-GPU call: bench at N ranks (headline weak + sharded strong record + e2e with the host copy ceiling), host topology
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2d_bench_n$N.json 2> gpurun_out/r2d_bench_n$N.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2d_bench_n$N.json; tail -5 gpurun_out/r2d_bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 tools/host_copy_ceiling.py > gpurun_out/r2d_host_n$N.json 2> gpurun_out/r2d_host_n$N.err; echo "host rc=$?"
cat gpurun_out/r2d_host_n$N.json | cut -c1-1500
