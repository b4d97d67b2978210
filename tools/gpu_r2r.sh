#!/bin/bash
# 2-GPU validation: the K-split test, then bench.py at N=2 (headline + sharded + ksplit + e2e with the host ceiling)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ksplit_gpu.py -m gpu -q -x > gpurun_out/r2r_pytest_ksplit.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2r_pytest_ksplit.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2r_bench_n2.json 2> gpurun_out/r2r_bench_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
j = json.load(open("gpurun_out/r2r_bench_n2.json"))
print("value", j["value"], "e2e", j["e2e"]["value"], j["e2e"]["host_copy_ceiling"])
print(json.dumps(j["sharded"], indent=1))
PY
tail -5 gpurun_out/r2r_bench_n2.err | cut -c1-300
