import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from einsums_b200 import ops, KF_NAMES
ops.init(0)
if len(sys.argv) > 1: ops.set_option('tcgen05_pair', int(sys.argv[1]))
for (M,N,K,beta) in ((40000, 10000, 400, 1.0), (40000, 10000, 400, 0.0), (8192, 8192, 512, 0.0), (4096,4096,768, 0.0), (4096, 4096, 4096, 0.0)):
    A = torch.rand(M, K, dtype=torch.float32, device="cuda")*2-1; B = torch.rand(N, K, dtype=torch.float32, device="cuda")*2-1
    C = torch.zeros(M, N, dtype=torch.float32, device="cuda")
    for _ in range(2): d = ops.contract("ij", C, "ik", A, "jk", B, 1.0, beta)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ops.contract("ij", C, "ik", A, "jk", B, 1.0, beta)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/5
    print(M,N,K,"beta",beta, ops.last_kernel(), "ms", round(ms,3), "TFLOP/s", round(2*M*N*K/ms/1e9,1), flush=True)
