#!/bin/bash
# round 2, third GPU call: parity suite; prefetch distance A/B on Wmbej / Wmnij; TMA-staged permute A/B over the 24-order sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2c_pytest.log
for d in 2 3; do echo "dist=$d"; EB200_TMA_DIST=$d python tools/probe_wmbej.py 40 160; EB200_TMA_DIST=$d python tools/probe_wmbej.py 60 240; EB200_TMA_DIST=$d python tools/probe.py wmnij; done > gpurun_out/r2c_dist.txt 2>&1; cat gpurun_out/r2c_dist.txt
EB200_PERMUTE_TMA=2 timeout 600 python tools/run_configs.py 4 > gpurun_out/r2c_permute_tma2.jsonl 2>&1
EB200_PERMUTE_TMA=0 timeout 600 python tools/run_configs.py 4 > gpurun_out/r2c_permute_tma0.jsonl 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/r2c_permute_tma2.jsonl", "gpurun_out/r2c_permute_tma0.jsonl"):
    print(f)
    for l in open(f):
        try: j = json.loads(l)
        except Exception: print(l[:200]); continue
        if "gbs_min" in j: print(" ", j["config"], j["gbs_min"], j["gbs_median"], j["gbs_max"], j["worst_order"])
PY
