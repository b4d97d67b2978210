#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-configs 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])"
