import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from einsums_b200 import ops, KF_NAMES
ops.init(0)
torch.manual_seed(0)
m, n, k = 256, 256, 64
for a_idx, b_idx in (("ik", "jk"), ("ki", "jk"), ("ik", "kj"), ("ki", "kj")):
    ext = dict(i=m, j=n, k=k)
    A = torch.rand(*[ext[x] for x in a_idx], dtype=torch.float32, device="cuda") * 2 - 1
    B = torch.rand(*[ext[x] for x in b_idx], dtype=torch.float32, device="cuda") * 2 - 1
    C = torch.zeros(m, n, dtype=torch.float32, device="cuda")
    d = ops.contract("ij", C, a_idx, A, b_idx, B)
    ref = torch.einsum(f"{a_idx},{b_idx}->ij", A.double(), B.double())
    err = (C.double() - ref).abs()
    print(a_idx, b_idx, KF_NAMES[d.kernel_used], "max rel err", (err.max() / ref.abs().max()).item())
    if False:
        bad = (err > 1e-3)
        rows = bad.any(1).nonzero().flatten(); cols = bad.any(0).nonzero().flatten()
        print("   bad rows", rows[:12].tolist(), "... count", rows.numel(), " bad cols", cols[:12].tolist(), "... count", cols.numel())
        # try to identify: compare with partial-k reference
        for kk in (8, 16, 32, 64):
            refp = torch.einsum(f"{a_idx},{b_idx}->ij", A.double().narrow(a_idx.index("k"), 0, kk), B.double().narrow(b_idx.index("k"), 0, kk))
            print("   vs first", kk, "k:", ((C.double() - refp).abs().max() / ref.abs().max()).item())
