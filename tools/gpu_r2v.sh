#!/bin/bash
# checkpoint: smoke, full parity suite, full bench line (N = 1), reference arm
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2v_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2v_pytest.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2v_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/r2v_bench_reference.json 2> gpurun_out/r2v_bench_reference.err; echo "ref rc=$?"
python - <<'PY'
import json
j = json.load(open("gpurun_out/r2v_bench.json"))
print("value", j["value"], "roofline", j["roofline"]["frac"], "traffic", j["roofline"]["traffic"], "e2e", j["e2e"]["value"], j["e2e"]["host_copy_ceiling"]["frac"], "cpu", j["cpu_baseline"]["value"])
print("sharded", j["sharded"]["tflops_total"], j["sharded"]["frac_of_dmma"])
for r in j["configs"]["records"]:
    print("  ", r["config"], "|", r["op"][:50], "|", r.get("tflops"), r.get("frac_of_fp64_peak"), r.get("gbs", r.get("gbs_min")), r.get("frac_of_hbm_copy", r.get("frac_min")), r.get("kernel", ""))
r = json.load(open("gpurun_out/r2v_bench_reference.json"))
print("reference", r["value"], r["cpu_baseline"]["cores"])
PY
