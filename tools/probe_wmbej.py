"""cfg3a Wmbej (generic fused-permutation contraction) at a chosen size; for ncu captures and timing."""
import sys, torch
sys.path.insert(0, ".")
from einsums_b200 import ops
ops.init(0)
if len(sys.argv) > 3: ops.set_option('tma_generic', int(sys.argv[3]))
o, v = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (40, 160)
t = torch.rand(o, o, v, v, dtype=torch.float64, device="cuda") - 0.5
g = torch.rand(o, o, v, v, dtype=torch.float64, device="cuda") - 0.5
W = torch.rand(o, v, v, o, dtype=torch.float64, device="cuda") - 0.5
fn = lambda: ops.contract("mbej", W, "jnfb", t, "mnef", g, -0.5, 1.0)
for _ in range(2): fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(3): d = fn()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"Wmbej o={o} v={v}: {ms:.3f} ms  {2.0*o**3*v**3/ms/1e9:.2f} TFLOP/s  kernel={ops.last_kernel()} launches={d.launches}")
