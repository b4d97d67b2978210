"""c128 multi-index contraction: TTGT around gemm_tma_z (tma_generic = 1) against the FMA-pipe tile kernel (tma_generic = 0)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from einsums_b200 import ops
ops.init(0)
def r(*s):
    return torch.complex(torch.rand(*s, dtype=torch.float64, device="cuda") - 0.5, torch.rand(*s, dtype=torch.float64, device="cuda") - 0.5)
for (o, v) in ((40, 160),):
    t, g, W = r(o, o, v, v), r(o, o, v, v), r(o, v, v, o)
    rec = {"op": f"Wmbej c128 o={o} v={v}"}
    outs = {}
    for mode in (1, 0):
        ops.set_option("tma_generic", mode)
        Ww = W.clone()
        for _ in range(2): d = ops.contract("mbej", Ww, "jnfb", t, "mnef", g, -0.5, 1.0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(3): d = ops.contract("mbej", Ww, "jnfb", t, "mnef", g, -0.5, 1.0)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        rec[ops.last_kernel()] = {"ms": round(ms, 3), "real_tflops": round(8.0 * o**3 * v**3 / ms / 1e9, 2), "launches": d.launches}
        Ww = W.clone()
        ops.contract("mbej", Ww, "jnfb", t, "mnef", g, -0.5, 1.0)
        outs[mode] = Ww
    ops.set_option("tma_generic", 1)
    rec["max_rel_diff_between_kernels"] = float((outs[0] - outs[1]).abs().max() / outs[0].abs().max())
    print(json.dumps(rec), flush=True)
