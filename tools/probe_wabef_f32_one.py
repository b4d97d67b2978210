import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from einsums_b200 import ops
ops.init(0)
beta = float(sys.argv[1])
M, N, K = 40000, 10000, 400
A = torch.rand(M, K, dtype=torch.float32, device="cuda") - 0.5
B = torch.rand(N, K, dtype=torch.float32, device="cuda") - 0.5
C = torch.zeros(M, N, dtype=torch.float32, device="cuda")
for _ in range(3): ops.contract("ij", C, "ik", A, "jk", B, 1.0, beta)
torch.cuda.synchronize()
print(ops.last_kernel())
