import sys, torch
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
N = 1 << 28
u = torch.rand(N, dtype=torch.float64, device="cuda"); w = torch.rand(N, dtype=torch.float64, device="cuda")
t = timeit(lambda: ops.scal(N, 1.0000001, u, 1)); print(f"dscal 2^28 (in-place RMW, 16 B/elem): {16.0*N/t/1e12:.2f} TB/s")
t = timeit(lambda: u.mul_(1.0000001)); print(f"torch mul_ in place: {16.0*N/t/1e12:.2f} TB/s")
t = timeit(lambda: torch.mul(u, 1.0000001, out=w)); print(f"torch mul out of place: {16.0*N/t/1e12:.2f} TB/s")
t = timeit(lambda: w.copy_(u)); print(f"torch copy: {16.0*N/t/1e12:.2f} TB/s")
t = timeit(lambda: ops.axpy(N, 0.5, u, 1, w, 1)); print(f"daxpy: {24.0*N/t/1e12:.2f} TB/s")
