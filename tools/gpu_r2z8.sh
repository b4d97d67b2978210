#!/bin/bash
N=8
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2z_bench_n$N.json 2> gpurun_out/r2z_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
for l in open("gpurun_out/r2z_bench_n8.json"):
    if l.startswith("{"):
        d=json.loads(l); sh=d.get("sharded",{}); print("value", d["value"], "e2e", d["e2e"]["value"], "ceiling frac", d["e2e"].get("host_copy_ceiling",{}).get("frac"), "sharded", sh.get("tflops_total"), sh.get("parity_rel_err"), "ksplit", sh.get("ksplit",{}).get("tflops_total"), sh.get("ksplit",{}).get("speedup_vs_single_gpu"))
PY
tail -2 gpurun_out/r2z_bench_n$N.err
