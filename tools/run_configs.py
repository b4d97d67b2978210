#!/usr/bin/env python
"""The five BASELINE.json configs at their full single-GPU sizes: one record per einsum / permute with the achieved TFLOP/s
or GB/s and the fraction of the measured roofline (FP64 DMMA peak from tools/fp64_peaks, HBM copy peak from
MEASURED_PEAKS.json). Every timed call is preceded by a parity spot check against torch (float64) on a slice.

Importable (bench.py puts these records into its JSON line as "configs") and runnable:
    python tools/run_configs.py [1 2 3 4]      -> one JSON line per record, then the peaks and the clocks seen during the run
"""
import itertools, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from einsums_b200 import ops, KF_NAMES


def peaks():
    """(HBM copy GB/s, its source), (FP64 DMMA TFLOP/s, its source)"""
    try:
        hbm, hsrc = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    except Exception:
        hbm, hsrc = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    try:
        out = subprocess.run([os.path.join(ROOT, "tools", "fp64_peaks")], capture_output=True, text=True, timeout=120).stdout
        fp64, fsrc = json.loads(out.strip().splitlines()[-1])["dmma_tflops"], "tools/fp64_peaks measured in this run (DMMA issue loop)"
    except Exception:
        fp64, fsrc = 148 * 64 * 2 * 1.965e9 / 1e12, "nominal 148 SM x 64 FMA/clk x 1.965 GHz"
    return hbm, hsrc, fp64, fsrc


def rnd(*s):
    return torch.rand(*s, dtype=torch.float64, device="cuda") * 2 - 1


def timeit(fn, warm=2, reps=5):
    """median of `reps` CUDA-event timings on the current stream; sub-millisecond ops are timed as 20 (or 4) back-to-back launches"""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    inner = 1 if e0.elapsed_time(e1) > 5 else (4 if e0.elapsed_time(e1) > 1 else 20)
    ts = []
    for _ in range(reps):
        e0.record()
        for _ in range(inner):
            fn()
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / inner)
    return sorted(ts)[len(ts) // 2]


class Recorder:
    def __init__(self, hbm, fp64, emit=None):
        self.hbm, self.fp64, self.records, self.emit = hbm, fp64, [], emit

    def __call__(self, cfg, op, ms, flops=None, nbytes=None, **kw):
        d = {"config": cfg, "op": op, "ms": round(ms, 4)}
        if flops:
            d.update(tflops=round(flops / ms / 1e9, 2), frac_of_fp64_peak=round(flops / ms / 1e9 / self.fp64, 3))
        if nbytes:
            d.update(gbs=round(nbytes / ms / 1e6, 1), frac_of_hbm_copy=round(nbytes / ms / 1e6 / self.hbm, 3))
        d.update(kw)
        self.add(d)

    def add(self, d):
        self.records.append(d)
        if self.emit:
            self.emit(d)


def cfg1(report):
    n = 100
    g, D, J, K = rnd(n, n, n, n), rnd(n, n), rnd(n, n), rnd(n, n)
    ops.contract("pq", J, "pqrs", g, "rs", D, 2.0, 0.0); ops.contract("pq", K, "prqs", g, "rs", D, -1.0, 0.0)
    errJ = ((J - 2 * torch.einsum("pqrs,rs->pq", g, D)).abs().max() / J.abs().max()).item()
    errK = ((K + torch.einsum("prqs,rs->pq", g, D)).abs().max() / K.abs().max()).item()
    report("cfg1 HF n=100", "J (gemv)", timeit(lambda: ops.contract("pq", J, "pqrs", g, "rs", D, 2.0, 0.0)), 2.0 * n**4, 8.0 * n**4, parity=errJ)
    report("cfg1 HF n=100", "K (generic)", timeit(lambda: ops.contract("pq", K, "prqs", g, "rs", D, -1.0, 0.0)), 2.0 * n**4, 8.0 * n**4, parity=errK)


def cfg2(report, wabef=True):
    o, v = 20, 200
    g, t, Wm = rnd(o, o, v, v), rnd(o, o, v, v), rnd(o, o, o, o)
    W0 = Wm.clone()
    d = ops.contract("mnij", Wm, "mnef", g, "ijef", t, 0.25, 1.0)
    ref = W0 + 0.25 * torch.einsum("mnef,ijef->mnij", g, t)
    err = ((Wm - ref).abs().max() / ref.abs().max()).item()
    report("cfg2 CCD o=20 v=200", "Wmnij", timeit(lambda: ops.contract("mnij", Wm, "mnef", g, "ijef", t, 0.25, 1.0)), 2.0 * o**4 * v**2,
           kernel=ops.last_kernel(), split_k=int(d.split_k), parity=err)
    if wabef:
        Wa = rnd(v, v, v, v)
        report("cfg2 CCD o=20 v=200", "Wabef", timeit(lambda: ops.contract("abef", Wa, "mnef", g, "mnab", t, 0.25, 1.0), 1, 3), 2.0 * v**4 * o**2,
               kernel=ops.last_kernel())
        del Wa


def cfg3(report, permuted_k=True):
    o, v = 60, 240
    t, g, W = rnd(o, o, v, v), rnd(o, o, v, v), rnd(o, v, v, o)
    W0 = W[:2].clone()
    d = ops.contract("mbej", W, "jnfb", t, "mnef", g, -0.5, 1.0)
    ref = W0 - 0.5 * torch.einsum("jnfb,mnef->mbej", t, g[:2])
    err = ((W[:2] - ref).abs().max() / ref.abs().max()).item()
    report("cfg3 o=60 v=240", "Wmbej (generic, fused permutation)", timeit(lambda: ops.contract("mbej", W, "jnfb", t, "mnef", g, -0.5, 1.0), 1, 3),
           2.0 * o**3 * v**3, kernel=ops.last_kernel(), family=KF_NAMES[d.kernel_used], parity=err)
    del t, g, W
    torch.cuda.empty_cache()
    n = 300
    g = torch.empty(n, n, n, n, dtype=torch.float64, device="cuda")
    for i in range(0, n, 30):
        g[i:i + 30] = rnd(30, n, n, n)
    D, K, J = rnd(n, n), rnd(n, n), rnd(n, n)
    d = ops.contract("pq", K, "prqs", g, "rs", D, -1.0, 0.0)
    ref = -torch.einsum("rqs,rs->q", g[7], D)
    err = ((K[7] - ref).abs().max() / ref.abs().max()).item()
    report("cfg3 n=300", "Fock K (generic)", timeit(lambda: ops.contract("pq", K, "prqs", g, "rs", D, -1.0, 0.0), 1, 3), 2.0 * n**4, 8.0 * n**4,
           family=KF_NAMES[d.kernel_used], parity=err)
    report("cfg3 n=300", "Fock J (gemv)", timeit(lambda: ops.contract("pq", J, "pqrs", g, "rs", D, 2.0, 0.0), 1, 3), 2.0 * n**4, 8.0 * n**4)
    if permuted_k:
        # permuted-K: permute g[p,r,q,s] -> gp[p,q,r,s] then GEMV; 64.8 GB more memory
        gp = torch.empty(n, n, n, n, dtype=torch.float64, device="cuda")
        tp = timeit(lambda: ops.permute("pqrs", gp, "prqs", g), 1, 3)
        Kp = torch.empty(n, n, dtype=torch.float64, device="cuda")
        tg = timeit(lambda: ops.contract("pq", Kp, "pqrs", gp, "rs", D, -1.0, 0.0), 1, 3)
        errp = ((Kp - K).abs().max() / K.abs().max()).item()
        report("cfg3 n=300", "permuted-K: permute", tp, None, 16.0 * n**4)
        report("cfg3 n=300", "permuted-K: gemv", tg, 2.0 * n**4, 8.0 * n**4, parity_vs_generic=errp)
        report("cfg3 n=300", "permuted-K: total (permute + gemv)", tp + tg, 2.0 * n**4, None)
        del gp
    del g
    torch.cuda.empty_cache()


def cfg4(report, sizes=(64, 96, 128, 160), per_order=True):
    for n in sizes:
        A, B = rnd(n, n, n, n), torch.empty(n, n, n, n, dtype=torch.float64, device="cuda")
        res = {}
        for order in itertools.permutations(range(4)):
            c = "".join("ijkl"[x] for x in order)
            ops.permute(c, B, "ijkl", A)
            if n == 64:
                assert torch.equal(B, A.permute(*order)), c
            res[c] = round(16.0 * n**4 / timeit(lambda: ops.permute(c, B, "ijkl", A), 1, 3) / 1e6, 1)
        vals = sorted(res.values())
        worst = min(res, key=res.get)
        rec = {"config": f"cfg4 permute n={n}", "op": "24 rank-4 orders, alpha=1 beta=0 (16 n^4 bytes)", "gbs_min": vals[0], "gbs_median": vals[12],
               "gbs_max": vals[-1], "frac_min": round(vals[0] / report.hbm, 3), "frac_median": round(vals[12] / report.hbm, 3), "worst_order": worst,
               "parity": "bit-exact vs torch.permute (n=64, all orders)" if n == 64 else "checked at n=64"}
        if per_order:
            rec["all"] = res
        report.add(rec)
        Bt = torch.rand(n, n, n, n, dtype=torch.float64, device="cuda")
        ms = timeit(lambda: ops.permute("lkji", Bt, "ijkl", A, 0.5, 2.0), 1, 3)
        report(f"cfg4 permute n={n}", "lkji alpha=0.5 beta=2", ms, None, 24.0 * n**4)
        del A, B, Bt
    torch.cuda.empty_cache()


def cfg_types(report):
    """The other value types of the flat seam ({s,c,z}gemm and the generic contraction) on the tensor cores: FP32 / c64 on tcgen05 with
    3xTF32 splitting (real flops against a third of the nominal dense TF32 rate, 1125 / 3 = 375 TFLOP/s), c128 on the FP64 tensor pipe."""
    TF32X3 = 1125.0 / 3
    o, v, n = 40, 160, 4096
    def rt(dt, *s):
        if dt.is_complex:
            rd = torch.float32 if dt == torch.complex64 else torch.float64
            return torch.complex(torch.rand(*s, dtype=rd, device="cuda") - 0.5, torch.rand(*s, dtype=rd, device="cuda") - 0.5)
        return torch.rand(*s, dtype=dt, device="cuda") - 0.5
    for dt, name, mul, tol in ((torch.float32, "f32", 2.0, 1e-5), (torch.complex64, "c64", 8.0, 1e-5), (torch.complex128, "c128", 8.0, 1e-12)):
        peak, pname = (report.fp64, "frac_of_fp64_peak") if dt == torch.complex128 else (TF32X3, "frac_of_3xtf32_peak")
        A, B, C = rt(dt, n, n), rt(dt, n, n), rt(dt, n, n)
        C0 = C[:64].clone()
        ops.contract("ij", C, "ik", A, "kj", B, 0.5, 1.0)
        ref = C0.to(torch.complex128 if dt.is_complex else torch.float64) + 0.5 * (A[:64].to(torch.complex128 if dt.is_complex else torch.float64) @ B.to(torch.complex128 if dt.is_complex else torch.float64))
        err = ((C[:64] - ref).abs().max() / ref.abs().max()).item()
        assert err <= tol, (name, err)
        ms = timeit(lambda: ops.contract("ij", C, "ik", A, "kj", B, 0.5, 1.0))
        report.add({"config": f"types {name}", "op": f"gemm {n}^3 (ik,kj) beta=1", "ms": round(ms, 4), "real_tflops": round(mul * n**3 / ms / 1e9, 1),
                    pname: round(mul * n**3 / ms / 1e9 / peak, 3), "kernel": ops.last_kernel(), "parity": err, "parity_tol": tol})
        del A, B, C
        t, g, W = rt(dt, o, o, v, v), rt(dt, o, o, v, v), rt(dt, o, v, v, o)
        W0 = W[:1].clone()
        d = ops.contract("mbej", W, "jnfb", t, "mnef", g, -0.5, 1.0)
        wide = torch.complex128 if dt.is_complex else torch.float64
        ref = W0.to(wide) - 0.5 * torch.einsum("jnfb,mnef->mbej", t.to(wide), g[:1].to(wide))
        err = ((W[:1] - ref).abs().max() / ref.abs().max()).item()
        assert err <= tol, (name, err)
        ms = timeit(lambda: ops.contract("mbej", W, "jnfb", t, "mnef", g, -0.5, 1.0), 1, 3)
        report.add({"config": f"types {name}", "op": f"Wmbej o={o} v={v} (generic)", "ms": round(ms, 4), "real_tflops": round(mul * o**3 * v**3 / ms / 1e9, 1),
                    pname: round(mul * o**3 * v**3 / ms / 1e9 / peak, 3), "kernel": ops.last_kernel(), "launches": int(d.launches), "parity": err,
                    "parity_tol": tol})
        del t, g, W
    torch.cuda.empty_cache()


def run_all(which=("1", "2", "3", "4"), emit=None, wabef=True, permuted_k=True, per_order=True):
    hbm, hsrc, fp64, fsrc = peaks()
    report = Recorder(hbm, fp64, emit)
    if "1" in which: cfg1(report)
    if "2" in which: cfg2(report, wabef)
    if "3" in which: cfg3(report, permuted_k)
    if "4" in which: cfg4(report, per_order=per_order)
    if "T" in which:
        try:
            cfg_types(report)
        except Exception as e:  # the headline line must still be printed
            report.add({"config": "types", "error": repr(e)[:300]})
    return report.records, {"fp64_dmma_tflops": fp64, "fp64_source": fsrc, "hbm_copy_gbs": hbm, "hbm_source": hsrc}


if __name__ == "__main__":
    from bench import ClockSampler
    ops.init(0); torch.cuda.set_device(0)
    sampler = ClockSampler(0); sampler.start()
    _, pk = run_all(sys.argv[1:] or ("1", "2", "3", "4"), emit=lambda d: print(json.dumps(d), flush=True))
    print(json.dumps({"peaks": pk, "clocks": sampler.stop()}))
