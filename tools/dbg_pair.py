"""Bring-up of the CTA-pair (cta_group::2) FP32 kernel: small to large products against torch (float64 reference)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from einsums_b200 import ops
ops.init(0)
ops.set_option("tcgen05_pair", 2)
torch.manual_seed(1)
for (M, N, K, beta, lay) in ((256, 128, 32, 0.0, "ik,jk"), (256, 128, 64, 0.0, "ik,jk"), (512, 256, 96, 1.0, "ik,jk"), (300, 200, 100, 0.5, "ik,jk"),
                             (1024, 1024, 1024, 0.0, "ik,jk"), (1024, 1024, 1024, 0.0, "ik,kj"), (1024, 1024, 1024, 0.0, "ki,jk"), (1024, 1024, 1024, 1.0, "ki,kj"),
                             (4096, 4096, 4096, 0.0, "ik,jk"), (40000, 10000, 400, 1.0, "ik,jk")):
    a_idx, b_idx = lay.split(",")
    ext = {"i": M, "j": N, "k": K}
    A = torch.rand(*[ext[c] for c in a_idx], dtype=torch.float32, device="cuda") - 0.5
    B = torch.rand(*[ext[c] for c in b_idx], dtype=torch.float32, device="cuda") - 0.5
    C0 = torch.rand(M, N, dtype=torch.float32, device="cuda")
    C = C0.clone()
    d = ops.contract("ij", C, a_idx, A, b_idx, B, 1.0, beta)
    torch.cuda.synchronize()
    if M * N * K < 2e10:
        ref = beta * C0.double() + torch.einsum(f"{a_idx},{b_idx}->ij", A.double(), B.double())
        err = ((C.double() - ref).abs().max() / ref.abs().max()).item()
    else:
        ref = beta * C0[:256].double() + torch.einsum(f"{a_idx},{b_idx}->ij", A[:256].double(), B.double())
        err = ((C[:256].double() - ref).abs().max() / ref.abs().max()).item()
    for _ in range(2): ops.contract("ij", C, a_idx, A, b_idx, B, 1.0, beta)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5): ops.contract("ij", C, a_idx, A, b_idx, B, 1.0, beta)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(M, N, K, lay, "beta", beta, ops.last_kernel(), "err", f"{err:.2e}", "ms", round(ms, 3), "TFLOP/s", round(2.0 * M * N * K / ms / 1e9, 1), flush=True)
