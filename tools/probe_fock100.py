"""cfg1 Fock J / K at n=100 (for ncu captures and quick timing)."""
import sys, torch
sys.path.insert(0, ".")
from einsums_b200 import ops
ops.init(0)
n = 100
g = torch.rand(n, n, n, n, dtype=torch.float64, device="cuda"); D = torch.rand(n, n, dtype=torch.float64, device="cuda")
J = torch.zeros(n, n, dtype=torch.float64, device="cuda"); K = torch.zeros(n, n, dtype=torch.float64, device="cuda")
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for name, fn in (("J", lambda: ops.contract("pq", J, "pqrs", g, "rs", D, 2.0, 0.0)), ("K", lambda: ops.contract("pq", K, "prqs", g, "rs", D, -1.0, 0.0))):
    ms = timeit(fn)
    print(name, f"{ms*1e3:.1f} us", f"{8.0*n**4/ms/1e6:.0f} GB/s", flush=True)
