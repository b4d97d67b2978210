"""Where does the generic FP32 instantiation lose against the plain kernel at the same M = N = K = 6400?"""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from einsums_b200 import ops
ops.init(0)
r = lambda *s: torch.rand(*s, dtype=torch.float32, device="cuda") - 0.5
def run(name, c, C, a, A, b, B, alpha, beta, flops):
    for _ in range(2): d = ops.contract(c, C, a, A, b, B, alpha, beta)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5): d = ops.contract(c, C, a, A, b, B, alpha, beta)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(json.dumps({"op": name, "kernel": ops.last_kernel(), "ms": round(ms, 3), "tflops": round(flops / ms / 1e9, 1)}), flush=True)
n = 6400
o, v = 40, 160
fl = 2.0 * n**3
run("plain ik,jk beta=1", "ij", r(n, n), "ik", r(n, n), "jk", r(n, n), -0.5, 1.0, fl)
run("plain ik,jk beta=0", "ij", r(n, n), "ik", r(n, n), "jk", r(n, n), -0.5, 0.0, fl)
run("plain ki,jk beta=1 (A M-major)", "ij", r(n, n), "ki", r(n, n), "jk", r(n, n), -0.5, 1.0, fl)
# generic machinery on a trivially multi-index problem: M = (a, b) split that does not merge in C
run("generic: C[b,a,j] <- A[a,b,k] B[j,k] (K-major both)", "baj", r(160, 40, n), "abk", r(40, 160, n), "jk", r(n, n), -0.5, 1.0, fl)
run("generic: same, beta=0", "baj", r(160, 40, n), "abk", r(40, 160, n), "jk", r(n, n), -0.5, 0.0, fl)
run("Wmbej beta=1", "mbej", r(o, v, v, o), "jnfb", r(o, o, v, v), "mnef", r(o, o, v, v), -0.5, 1.0, fl)
run("Wmbej beta=0", "mbej", r(o, v, v, o), "jnfb", r(o, o, v, v), "mnef", r(o, o, v, v), -0.5, 0.0, fl)
run("Wmbej-like, C in GEMM order [j,b,m,e]", "jbme", r(o, v, o, v), "jnfb", r(o, o, v, v), "mnef", r(o, o, v, v), -0.5, 1.0, fl)
