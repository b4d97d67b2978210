"""c64 contractions: gemm_tf32x3_tcgen05<c64> (packed operands, real 3xTF32 product) against the FMA-pipe tile kernel (tcgen05_generic = 0)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from einsums_b200 import ops
ops.init(0)
if len(sys.argv) > 1: ops.set_option('tcgen05_pair', int(sys.argv[1]))

def run(name, c, C, a, A, b, B, alpha, beta, flops):
    rec = {"op": name}
    outs = {}
    for mode in (1, 0):
        ops.set_option("tcgen05_generic", mode)
        Cw = C.clone()
        for _ in range(2): d = ops.contract(c, Cw, a, A, b, B, alpha, beta)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(5): d = ops.contract(c, Cw, a, A, b, B, alpha, beta)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        rec[ops.last_kernel()] = {"ms": round(ms, 3), "real_tflops": round(flops / ms / 1e9, 1), "launches": d.launches}
        Cw = C.clone()
        ops.contract(c, Cw, a, A, b, B, alpha, beta)
        outs[mode] = Cw
    ops.set_option("tcgen05_generic", 1)
    rec["max_rel_diff_between_kernels"] = float((outs[0] - outs[1]).abs().max() / outs[0].abs().max())
    print(json.dumps(rec), flush=True)

def r(*s):
    return torch.complex(torch.rand(*s, dtype=torch.float32, device="cuda") - 0.5, torch.rand(*s, dtype=torch.float32, device="cuda") - 0.5)
n = 4096
run("cgemm 4096^3 nn", "ij", r(n, n), "ik", r(n, n), "kj", r(n, n), 1.0, 0.0, 8.0 * n**3)
run("cgemm 4096^3 nt beta=1", "ij", r(n, n), "ik", r(n, n), "jk", r(n, n), 0.5 + 0.5j, 1.0, 8.0 * n**3)
run("cgemm 8192x8192x512", "ij", r(8192, 8192), "ik", r(8192, 512), "jk", r(8192, 512), 1.0, 1.0, 8.0 * 8192 * 8192 * 512)
for (o, v) in ((40, 160),):
    run(f"Wmbej c64 o={o} v={v}", "mbej", r(o, v, v, o), "jnfb", r(o, o, v, v), "mnef", r(o, o, v, v), -0.5, 1.0, 8.0 * o**3 * v**3)
run("batched c64 24 x 1024x1024x512", "xyij", r(4, 6, 1024, 1024), "xyik", r(4, 6, 1024, 512), "yxkj", r(6, 4, 512, 1024), 1.0, 0.0, 8.0 * 24 * 1024 * 1024 * 512)
