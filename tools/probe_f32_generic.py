"""FP32 multi-index contractions: gemm_tf32x3_tcgen05<generic> against the FMA-pipe tile kernel (tcgen05_generic = 0) on the same inputs."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from einsums_b200 import ops, KF_NAMES
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
        for _ in range(5): ops.contract(c, Cw, a, A, b, B, alpha, beta)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        rec[ops.last_kernel()] = {"ms": round(ms, 3), "tflops": round(flops / ms / 1e9, 1)}
        Cw = C.clone()
        ops.contract(c, Cw, a, A, b, B, alpha, beta)
        outs[mode] = Cw
    ops.set_option("tcgen05_generic", 1)
    rec["max_rel_diff_between_kernels"] = float((outs[0] - outs[1]).abs().max() / outs[0].abs().max())
    print(json.dumps(rec), flush=True)

r = lambda *s: torch.rand(*s, dtype=torch.float32, device="cuda") - 0.5
for (o, v) in ((40, 160), (60, 240)):
    run(f"Wmbej f32 o={o} v={v}", "mbej", r(o, v, v, o), "jnfb", r(o, o, v, v), "mnef", r(o, o, v, v), -0.5, 1.0, 2.0 * o**3 * v**3)
o, v = 20, 200
run("Wabef f32 o=20 v=200 (nmef order)", "abef", r(v, v, v, v), "nmef", r(o, o, v, v), "mnab", r(o, o, v, v), 0.25, 1.0, 2.0 * o**2 * v**4)
run("batched 2 idx f32", "xyij", r(4, 6, 1024, 1024), "xyik", r(4, 6, 1024, 512), "yxkj", r(6, 4, 512, 1024), 1.0, 0.0, 2.0 * 24 * 1024 * 1024 * 512)
o, v = 41, 163
run(f"Wmbej f32 o={o} v={v} (odd extents: packed)", "mbej", r(o, v, v, o), "jnfb", r(o, o, v, v), "mnef", r(o, o, v, v), -0.5, 1.0, 2.0 * o**3 * v**3)
