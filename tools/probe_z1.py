"""one c128 product for ncu captures"""
import sys, torch
sys.path.insert(0, ".")
from einsums_b200 import ops
ops.init(0)
M = N = K = 4096
A = torch.randn(M, K, dtype=torch.complex128, device="cuda"); B = torch.randn(N, K, dtype=torch.complex128, device="cuda"); C = torch.zeros(M, N, dtype=torch.complex128, device="cuda")
for _ in range(3): ops.contract("ij", C, "ik", A, "jk", B, 1.0, 0.0)
torch.cuda.synchronize(); print(ops.last_kernel())
