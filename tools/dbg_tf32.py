import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from einsums_b200 import ops, KF_NAMES
ops.init(0)
torch.manual_seed(0)
for (M, N, K) in ((256, 256, 64), (256, 384, 512), (1000, 520, 260), (512, 512, 1100), (512, 384, 16384), (400, 400, 40000), (4096, 4096, 4096)):
    A = torch.rand(M, K, dtype=torch.float32, device="cuda") * 2 - 1
    B = torch.rand(N, K, dtype=torch.float32, device="cuda") * 2 - 1
    C0 = torch.rand(M, N, dtype=torch.float32, device="cuda") * 2 - 1
    ref = (0.5 * (A.double() @ B.double().T) + 2.0 * C0.double())
    C = C0.clone()
    d = ops.contract("ij", C, "ik", A, "jk", B, 0.5, 2.0)
    torch.cuda.synchronize()
    err = ((C.double() - ref).abs().max() / ref.abs().max()).item()
    os.environ["X"] = "1"
    print(M, N, K, "kernel", KF_NAMES[d.kernel_used], "launches", d.launches, "rel err", err, flush=True)
    if M == 4096:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(2): ops.contract("ij", C, "ik", A, "jk", B, 1.0, 0.0)
        e0.record()
        for _ in range(5): ops.contract("ij", C, "ik", A, "jk", B, 1.0, 0.0)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print("  4096^3: ms", ms, "TFLOP/s", 2 * 4096**3 / ms / 1e9)
