import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from einsums_b200 import ops
ops.init(0)
ops.set_option("tcgen05_pair", int(sys.argv[1]) if len(sys.argv) > 1 else 2)
n = 4096
A = torch.rand(n, n, dtype=torch.float32, device="cuda") - 0.5
B = torch.rand(n, n, dtype=torch.float32, device="cuda") - 0.5
C = torch.zeros(n, n, dtype=torch.float32, device="cuda")
for _ in range(3): ops.contract("ij", C, "ik", A, "jk", B, 1.0, 0.0)
torch.cuda.synchronize()
print(ops.last_kernel())
