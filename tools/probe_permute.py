"""One rank-4 permute order at one size, a few launches (for ncu captures and quick timing): probe_permute.py <order> <n> [beta]"""
import sys, torch
sys.path.insert(0, ".")
from einsums_b200 import ops
ops.init(0)
order, n = sys.argv[1], int(sys.argv[2])
beta = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
A = torch.rand(n, n, n, n, dtype=torch.float64, device="cuda")
B = torch.zeros(n, n, n, n, dtype=torch.float64, device="cuda")
fn = lambda: ops.permute(order, B, "ijkl", A, 1.0, beta)
for _ in range(3): fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(10): fn()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"permute {order}<-ijkl n={n} beta={beta}: {ms:.4f} ms  {16.0*n**4/ms/1e6:.1f} GB/s")
