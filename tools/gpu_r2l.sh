#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2l_pytest.log | cut -c1-300
python tools/probe_level12.py > gpurun_out/r2l_level12.jsonl 2>&1; cat gpurun_out/r2l_level12.jsonl
# ncu launch list of the bench command (per-launch times are cold-cache and serialised: only the kernel's SHARE of the step is meaningful)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs --no-sharded --no-e2e > gpurun_out/r2l_bench_under_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches_r2.csv")) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; kn = h.index("Kernel Name"); mv = h.index("Metric Value")
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[hdr + 1:]:
    try: v = float(r[mv].replace(",", ""))
    except ValueError: continue
    name = r[kn].split("(")[0][:90]; tot[name] += v; cnt[name] += 1
s = sum(tot.values())
with open("gpurun_out/launches_r2_summary.txt", "w") as f:
    for k, v in tot.most_common():
        line = f"{v/1e6:12.3f} ms  {100*v/s:5.1f} %  x{cnt[k]:<4d} {k}"
        print(line); f.write(line + "\n")
PY
