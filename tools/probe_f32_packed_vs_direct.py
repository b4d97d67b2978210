"""FP32 Wmbej: direct generic kernel (tcgen05_generic = 1) against packed operands (3) on the same inputs."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from einsums_b200 import ops
ops.init(0)
r = lambda *s: torch.rand(*s, dtype=torch.float32, device="cuda") - 0.5
for (o, v) in ((40, 160), (60, 240), (20, 200)):
    t, g, W = r(o, o, v, v), r(o, o, v, v), r(o, v, v, o)
    for mode in (1, 3):
        ops.set_option("tcgen05_generic", mode)
        for _ in range(2): d = ops.contract("mbej", W, "jnfb", t, "mnef", g, -0.5, 1.0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(5): d = ops.contract("mbej", W, "jnfb", t, "mnef", g, -0.5, 1.0)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(json.dumps({"op": f"Wmbej f32 o={o} v={v}", "mode": mode, "kernel": ops.last_kernel(), "ms": round(ms, 3), "tflops": round(2.0 * o**3 * v**3 / ms / 1e9, 1), "launches": d.launches}), flush=True)
