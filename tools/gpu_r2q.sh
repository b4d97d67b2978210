#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2q_pytest.log | cut -c1-300
python tools/probe.py wmnij gemm 2>&1 | tail -6
python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
from einsums_b200 import ops
ops.init(0)
for (M, N, K) in ((1600, 1600, 160000), (800, 1600, 80000), (200, 1600, 20000), (3000, 3000, 3000), (2048, 2048, 2048), (5000, 1000, 4000)):
    A = torch.rand(M, K, dtype=torch.float64, device="cuda"); B = torch.rand(N, K, dtype=torch.float64, device="cuda"); C = torch.zeros(M, N, dtype=torch.float64, device="cuda")
    fn = lambda: ops.contract("ij", C, "ik", A, "jk", B)
    for _ in range(2): d = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(3): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{M}x{N}x{K}: {ms:.3f} ms {2.0*M*N*K/ms/1e9:.2f} TFLOP/s split={d.split_k} kernel={ops.last_kernel()}", flush=True)
PY
