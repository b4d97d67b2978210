"""Quick device probe of the main kernels (development aid; bench.py is the contract)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from einsums_b200 import ops, KF_NAMES

ops.init(0)
torch.cuda.set_device(0)

def timeit(fn, reps=5, warm=2, inner=None):
    """median over `reps` of the per-call device time; `inner` calls are enqueued back to back between the two events so that
    host launch latency (~20-30 us through Python/ctypes) does not pollute sub-millisecond kernels."""
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    if inner is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        one = e0.elapsed_time(e1)
        inner = 1 if one > 5.0 else (4 if one > 1.0 else 20)
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner): fn()
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / inner)
    ts.sort()
    return ts[len(ts)//2]

def rnd(*s): return torch.rand(*s, dtype=torch.float64, device="cuda") * 2 - 1

which = sys.argv[1:] or ["wabef", "wmnij", "wmbej", "fock", "permute", "gemm"]
out = {}
if "wabef" in which:
    o, v = 20, int(os.environ.get("V", 200))
    g, t, W = rnd(o, o, v, v), rnd(o, o, v, v), rnd(v, v, v, v)
    d = [None]
    def f(): d[0] = ops.contract("abef", W, "mnef", g, "mnab", t, 0.25, 1.0)
    ms = timeit(f, 5, 2)
    fl = 2.0 * v**4 * o**2
    print(json.dumps({"op": "Wabef", "o": o, "v": v, "ms": ms, "tflops": fl / ms / 1e9, "kernel": KF_NAMES[d[0].kernel_used], "split": d[0].split_k}), flush=True)
    del W
if "wmnij" in which:
    o, v = 20, 200
    g, t, W = rnd(o, o, v, v), rnd(o, o, v, v), rnd(o, o, o, o)
    d = [None]
    def f(): d[0] = ops.contract("mnij", W, "mnef", g, "ijef", t, 0.25, 1.0)
    ms = timeit(f, 10, 3)
    fl = 2.0 * o**4 * v**2
    print(json.dumps({"op": "Wmnij", "ms": ms, "tflops": fl / ms / 1e9, "kernel": KF_NAMES[d[0].kernel_used], "split": d[0].split_k}), flush=True)
if "wmbej" in which:
    o, v = int(os.environ.get("O", 30)), int(os.environ.get("VV", 120))
    t, g, W = rnd(o, o, v, v), rnd(o, o, v, v), rnd(o, v, v, o)
    d = [None]
    def f(): d[0] = ops.contract("mbej", W, "jnfb", t, "mnef", g, -0.5, 1.0)
    ms = timeit(f, 5, 2)
    fl = 2.0 * o**3 * v**3
    print(json.dumps({"op": "Wmbej", "o": o, "v": v, "ms": ms, "tflops": fl / ms / 1e9, "kernel": KF_NAMES[d[0].kernel_used], "split": d[0].split_k}), flush=True)
if "fock" in which:
    for n in (100, 200):
        g, D, F = rnd(n, n, n, n), rnd(n, n), rnd(n, n)
        d = [None]
        def fj(): d[0] = ops.contract("pq", F, "pqrs", g, "rs", D, 2.0, 0.0)
        ms = timeit(fj, 10, 3)
        print(json.dumps({"op": "FockJ", "n": n, "ms": ms, "gbs": 8.0 * n**4 / ms / 1e6, "kernel": KF_NAMES[d[0].kernel_used], "split": d[0].split_k}), flush=True)
        def fk(): d[0] = ops.contract("pq", F, "prqs", g, "rs", D, -1.0, 1.0)
        ms = timeit(fk, 10, 3)
        print(json.dumps({"op": "FockK", "n": n, "ms": ms, "gbs": 8.0 * n**4 / ms / 1e6, "kernel": KF_NAMES[d[0].kernel_used], "split": d[0].split_k}), flush=True)
        del g
if "permute" in which:
    import itertools
    for n in (64, 128):
        A, B = rnd(n, n, n, n), torch.empty(n, n, n, n, dtype=torch.float64, device="cuda")
        res = {}
        for order in itertools.permutations(range(4)):
            c = "".join("ijkl"[x] for x in order)
            ms = timeit(lambda: ops.permute(c, B, "ijkl", A), 5, 2)
            res[c] = round(16.0 * n**4 / ms / 1e6, 1)
        print(json.dumps({"op": "permute", "n": n, "gbs": res, "min": min(res.values()), "max": max(res.values())}), flush=True)
        del A, B
if "gemm" in which:
    for (m, n, k) in ((8192, 8192, 8192), (4096, 4096, 4096)):
        A, B, C = rnd(m, k), rnd(k, n), rnd(m, n)
        ms = timeit(lambda: ops.contract("ij", C, "ik", A, "kj", B, 1.0, 0.0), 3, 1)
        print(json.dumps({"op": "gemm_nn", "mnk": [m, n, k], "ms": ms, "tflops": 2.0 * m * n * k / ms / 1e9}), flush=True)
        ms = timeit(lambda: ops.contract("ij", C, "ik", A, "jk", B.t().contiguous(), 1.0, 0.0), 3, 1)
        print(json.dumps({"op": "gemm_nt", "mnk": [m, n, k], "ms": ms, "tflops": 2.0 * m * n * k / ms / 1e9}), flush=True)
