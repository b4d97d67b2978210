"""One contraction for ncu captures: probe_one.py f32wmbej | c64gemm | c64wmbej | zwmbej"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from einsums_b200 import ops
ops.init(0)
what = sys.argv[1]
rf = lambda *s: torch.rand(*s, dtype=torch.float32, device="cuda") - 0.5
rc = lambda *s: torch.complex(rf(*s), rf(*s))
rz = lambda *s: rc(*s).to(torch.complex128)
o, v = 40, 160
if what == "f32wmbej":
    W, t, g = rf(o, v, v, o), rf(o, o, v, v), rf(o, o, v, v)
    for _ in range(3): ops.contract("mbej", W, "jnfb", t, "mnef", g, -0.5, 1.0)
elif what == "c64gemm":
    n = 4096
    A, B, C = rc(n, n), rc(n, n), rc(n, n)
    for _ in range(3): ops.contract("ij", C, "ik", A, "kj", B, 1.0, 1.0)
elif what == "c64wmbej":
    W, t, g = rc(o, v, v, o), rc(o, o, v, v), rc(o, o, v, v)
    for _ in range(3): ops.contract("mbej", W, "jnfb", t, "mnef", g, -0.5, 1.0)
elif what == "zwmbej":
    W, t, g = rz(o, v, v, o), rz(o, o, v, v), rz(o, o, v, v)
    for _ in range(3): ops.contract("mbej", W, "jnfb", t, "mnef", g, -0.5, 1.0)
torch.cuda.synchronize()
print(what, ops.last_kernel())
