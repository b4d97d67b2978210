#!/bin/bash
# final validation at N = 2: bench both arms the way the driver launches them, K-split GPU tests
N=2
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2z_bench_n$N.json 2> gpurun_out/r2z_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
for l in open("gpurun_out/r2z_bench_n2.json"):
    if l.startswith("{"):
        d=json.loads(l); print("value", d["value"], "e2e", d["e2e"]["value"], "sharded", d.get("sharded",{}).get("tflops_total"), d.get("sharded",{}).get("parity_rel_err"), "ksplit", json.dumps(d.get("ksplit",{}))[:300])
PY
tail -3 gpurun_out/r2z_bench_n$N.err
timeout 300 python -m pytest tests -m gpu -q -x -k "ksplit" 2>&1 | tail -3
