#!/bin/bash
for KEEP in 4096; do export KEEP; echo keep=$KEEP
python - <<PY
import sys, torch
sys.path.insert(0, ".")
from einsums_b200 import ops
ops.init(0)
import os
ops.set_option("scratch_keep_mb", int(os.environ.get("KEEP", "4096")))
o, v = 60, 240
rf = lambda *s: torch.rand(*s, dtype=torch.float32, device="cuda") - 0.5
rc = lambda *s: torch.complex(rf(*s), rf(*s))
t, g, W = rc(o, o, v, v), rc(o, o, v, v), rc(o, v, v, o)
W0 = W[:1].clone()
for _ in range(2): d = ops.contract("mbej", W, "jnfb", t, "mnef", g, -0.5, 1.0)
torch.cuda.synchronize()
W.copy_(torch.zeros_like(W)); W[:1] = W0
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record(); d = ops.contract("mbej", W, "jnfb", t, "mnef", g, -0.5, 1.0); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
ref = W0.to(torch.complex128) - 0.5 * torch.einsum("jnfb,mnef->mbej", t.to(torch.complex128), g[:1].to(torch.complex128))
err = ((W[:1] - ref).abs().max() / ref.abs().max()).item()
print(f"Wmbej c64 o={o} v={v}: {ms:.2f} ms  {8.0*o**3*v**3/ms/1e9:.1f} real TFLOP/s  kernel={ops.last_kernel()} launches={d.launches} err={err:.2e}")
PY
done

timeout 300 python -m pytest tests -m gpu -q -x -k "packed or c64 or c128 or tcgen05 or ksplit or cabi" 2>&1 | tail -3
