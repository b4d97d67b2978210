#!/bin/bash
# round-2 closing validation on one GPU: smoke, the whole GPU suite, both bench arms the way the driver runs them
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2z_smoke.log | cut -c1-400
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2z_pytest.log
python bench.py --impl reference > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err; echo "ref rc=$?"
( time python bench.py ) > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2z_bench.err
python - <<PY
import json
for l in open("gpurun_out/r2z_bench.json"):
    if l.startswith("{"):
        d=json.loads(l); print("value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "sharded", d["sharded"]["tflops_total"], "cpu", d["cpu_baseline"]["value"])
        for r in d["configs"]["records"]:
            r.pop("all", None); print(json.dumps(r)[:260])
PY
