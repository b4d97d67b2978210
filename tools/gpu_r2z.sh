#!/bin/bash
mkdir -p gpurun_out
( time python bench.py ) > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2z_bench.err
python - <<PY
import json
for l in open("gpurun_out/r2z_bench.json"):
    if l.startswith("{"):
        d=json.loads(l); print(d["value"], d["roofline"]["frac"], d["e2e"]["value"])
        for r in d["configs"]["records"]:
            if r["config"].startswith("types"): print(json.dumps(r))
PY
