"""main-loop rate of the two contract_tma_f64 tile shapes against gemm_tma_f64 on a large product"""
import sys, torch
sys.path.insert(0, ".")
from einsums_b200 import ops
ops.init(0)
M = N = 8000; K = 8192
A = torch.rand(M, K, dtype=torch.float64, device="cuda"); B = torch.rand(N, K, dtype=torch.float64, device="cuda"); C = torch.zeros(M, N, dtype=torch.float64, device="cuda")
def run(label):
    fn = lambda: ops.contract("ij", C, "ik", A, "jk", B, 1.0, 0.0, split_k=(2 if "generic" in label else 0))
    for _ in range(2): d = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(3): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{label}: {ms:.2f} ms {2.0*M*N*K/ms/1e9:.2f} TFLOP/s kernel={ops.last_kernel()} split={d.split_k}", flush=True)
run("gemm_tma_f64")
ops.set_option("tma_tile", 0); run("generic 128x64 (split 2)")
ops.set_option("tma_tile", 1); run("generic 80x80 (split 2)")
