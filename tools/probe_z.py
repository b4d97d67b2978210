"""c128 products: real-flop rate (8 flops per complex multiply-add) of the tensor-pipe kernel (gemm_tma_z) and of the FMA-pipe tile kernel."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from einsums_b200 import ops, KF
ops.init(0)
for (M, N, K, a_idx, b_idx) in ((4096, 4096, 4096, "ik", "jk"), (4096, 4096, 4096, "ik", "kj"), (4096, 4096, 4096, "ki", "kj"), (8192, 8192, 512, "ik", "jk"), (20000, 20000, 200, "ki", "kj")):
    ext = dict(i=M, j=N, k=K)
    A = torch.randn(*[ext[x] for x in a_idx], dtype=torch.complex128, device="cuda")
    B = torch.randn(*[ext[x] for x in b_idx], dtype=torch.complex128, device="cuda")
    C = torch.zeros(M, N, dtype=torch.complex128, device="cuda")
    for force, name in ((0, "auto"), (KF["TILE_SIMT"], "forced tile_simt")):
        fn = lambda: ops.contract("ij", C, a_idx, A, b_idx, B, 1.0, 0.0, force_kernel=force)
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"zgemm {M}x{N}x{K} {a_idx},{b_idx} {name}: {ms:.3f} ms  {8.0*M*N*K/ms/1e9:.2f} real TFLOP/s  kernel={ops.last_kernel()}", flush=True)
    del A, B, C
