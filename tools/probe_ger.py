"""dger at a HBM-sized matrix (for ncu captures): A(m x m) += alpha x y^T through the flat seam."""
import sys, torch
sys.path.insert(0, ".")
from einsums_b200 import ops
ops.init(0)
m = 16384
A = torch.rand(m, m, dtype=torch.float64, device="cuda"); x = torch.rand(m, dtype=torch.float64, device="cuda")
for _ in range(4): ops.ger(m, m, 0.5, x, 1, x, 1, A, m)
torch.cuda.synchronize()
