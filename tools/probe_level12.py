"""Bandwidth of the flat-seam level-1/2 entry points (column-major dgemv N/T, ddot, daxpy, dger, ddirprod) at HBM-sized operands."""
import json, sys, torch
sys.path.insert(0, ".")
from einsums_b200 import ops
ops.init(0)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3
n = 16384
A = torch.rand(n, n, dtype=torch.float64, device="cuda"); x = torch.rand(n, dtype=torch.float64, device="cuda"); y = torch.zeros(n, dtype=torch.float64, device="cuda")
for tr in "nt":
    t = timeit(lambda: ops.gemv(tr, n, n, 1.0, A, n, x, 1, 0.0, y, 1))
    print(json.dumps({"op": f"dgemv {tr} {n}x{n}", "ms": round(t * 1e3, 3), "gbs": round(8.0 * n * n / t / 1e9, 1)}), flush=True)
N = 1 << 28
u = torch.rand(N, dtype=torch.float64, device="cuda"); v = torch.rand(N, dtype=torch.float64, device="cuda"); w = torch.zeros(N, dtype=torch.float64, device="cuda")
t = timeit(lambda: ops.dot(N, u, 1, v, 1)); print(json.dumps({"op": "ddot 2^28 (host result)", "ms": round(t * 1e3, 3), "gbs": round(16.0 * N / t / 1e9, 1)}), flush=True)
t = timeit(lambda: ops.axpy(N, 0.5, u, 1, v, 1)); print(json.dumps({"op": "daxpy 2^28", "ms": round(t * 1e3, 3), "gbs": round(24.0 * N / t / 1e9, 1)}), flush=True)
m = 16384
t = timeit(lambda: ops.ger(m, m, 0.5, x, 1, x, 1, A, m)); print(json.dumps({"op": f"dger {m}x{m}", "ms": round(t * 1e3, 3), "gbs": round(16.0 * m * m / t / 1e9, 1)}), flush=True)
Bm = torch.rand(m, m, dtype=torch.float64, device="cuda"); Cm = torch.zeros(m, m, dtype=torch.float64, device="cuda")
t = timeit(lambda: ops.contract("ij", Cm, "ij", A, "ij", Bm, 1.0, 0.0)); print(json.dumps({"op": f"direct product {m}x{m} (beta=0)", "ms": round(t * 1e3, 3), "gbs": round(24.0 * m * m / t / 1e9, 1)}), flush=True)
t = timeit(lambda: ops.contract("ij", Cm, "ij", A, "ij", Bm, 0.5, 2.0)); print(json.dumps({"op": f"direct product {m}x{m} (beta=2)", "ms": round(t * 1e3, 3), "gbs": round(32.0 * m * m / t / 1e9, 1)}), flush=True)
t = timeit(lambda: ops.scal(N, 1.0000001, u, 1)); print(json.dumps({"op": "dscal 2^28 (in place)", "ms": round(t * 1e3, 3), "gbs": round(16.0 * N / t / 1e9, 1)}), flush=True)
t = timeit(lambda: ops.copy(N, u, 1, w, 1)); print(json.dumps({"op": "dcopy 2^28", "ms": round(t * 1e3, 3), "gbs": round(16.0 * N / t / 1e9, 1)}), flush=True)
