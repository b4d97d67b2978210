"""FMA-pipe tile kernel throughput (forced), FP32 / FP64 / complex, plain GEMM and a fused-permutation contraction."""
import sys, torch
sys.path.insert(0, ".")
from einsums_b200 import ops, _lib
ops.init(0)
KF = {v: k for k, v in enumerate(_lib.KF_NAMES)}
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3
for dt, mul in ((torch.float32, 2), (torch.float64, 2), (torch.complex64, 8), (torch.complex128, 8)):
    n = 2048
    A = torch.randn(n, n, device="cuda", dtype=dt); B = torch.randn(n, n, device="cuda", dtype=dt); C = torch.zeros(n, n, device="cuda", dtype=dt)
    for spec in (("ij", "ik", "kj"), ("ij", "ki", "jk")):
        t = timeit(lambda: ops.contract(spec[0], C, spec[1], A, spec[2], B, 1.0, 0.0, force_kernel=KF["TILE_SIMT"]))
        print(f"{str(dt):18s} {spec}  {mul * n**3 / t / 1e12:7.2f} TFLOP/s", flush=True)
    o, v = 24, 96
    W = torch.zeros(o, v, v, o, device="cuda", dtype=dt); t2 = torch.randn(o, o, v, v, device="cuda", dtype=dt); g = torch.randn(o, o, v, v, device="cuda", dtype=dt)
    t = timeit(lambda: ops.contract("mbej", W, "jnfb", t2.permute(0, 1, 2, 3), "mnef", g, 1.0, 0.0, force_kernel=KF["TILE_SIMT"]))
    print(f"{str(dt):18s} Wmbej o={o} v={v}  {mul * o * o * v * v * o * v / t / 1e12:7.2f} TFLOP/s", flush=True)
