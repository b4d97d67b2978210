"""FP32 4096^3 in the four operand layouts, pair and single-CTA kernels (beta = 0)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from einsums_b200 import ops
ops.init(0)
n = 4096
r = lambda *s: torch.rand(*s, dtype=torch.float32, device="cuda") - 0.5
A, B, C = r(n, n), r(n, n), r(n, n)
for pair in (1, 0):
    ops.set_option("tcgen05_pair", pair)
    for a_idx, b_idx in (("ik", "jk"), ("ik", "kj"), ("ki", "jk"), ("ki", "kj")):
        for _ in range(2): ops.contract("ij", C, a_idx, A, b_idx, B, 1.0, 0.0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10): ops.contract("ij", C, a_idx, A, b_idx, B, 1.0, 0.0)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(json.dumps({"layout": f"{a_idx},{b_idx}", "kernel": ops.last_kernel(), "ms": round(ms, 4), "tflops": round(2.0 * n**3 / ms / 1e9, 1)}), flush=True)
