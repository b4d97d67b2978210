"""Block-diagonal contraction C_b = A_b * B_b: the reference's shape of execution (one einsum per block,
host wait after each, BlockAlgebra.hpp:36-47) against eb200_contract_grouped (equal blocks folded into one batched launch,
ragged blocks enqueued without waits). FP64, blocks packed in one allocation."""
import json, sys, torch
sys.path.insert(0, ".")
from einsums_b200 import ops
ops.init(0)

def views(flat, sizes):
    out, off = [], 0
    for n in sizes:
        out.append(flat[off:off + n * n].view(n, n)); off += n * n
    return out

def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3

for name, sizes in (("8 x 512", [512] * 8), ("64 x 128", [128] * 64), ("512 x 32", [32] * 512), ("8 x 2048", [2048] * 8),
                    ("ragged D2h-like", [700, 650, 300, 280, 120, 100, 60, 40]), ("ragged large", [3000, 2500, 1200, 900, 400, 300, 150, 80])):
    tot = sum(n * n for n in sizes)
    A = torch.rand(tot, dtype=torch.float64, device="cuda") - 0.5
    B = torch.rand(tot, dtype=torch.float64, device="cuda") - 0.5
    C = torch.zeros(tot, dtype=torch.float64, device="cuda")
    Av, Bv, Cv = views(A, sizes), views(B, sizes), views(C, sizes)
    flops = sum(2.0 * n**3 for n in sizes)
    group = ops.make_group("ij", Cv, "ik", Av, "kj", Bv)  # descriptors built once: the timed loops are C-ABI calls only
    singles = [ops.make_group("ij", [c], "ik", [a], "kj", [b]) for a, b, c in zip(Av, Bv, Cv)]
    stream = torch.cuda.current_stream()
    def loop_sync():
        for g in singles:
            ops.run_group(g)
            stream.synchronize()
    def grouped():
        ops.run_group(group)
    t0, t1 = timeit(loop_sync), timeit(grouped)
    groups = ops.run_group(group)
    ref = torch.cat([(a @ b).reshape(-1) for a, b in zip(Av, Bv)])
    err = float((C - ref).abs().max() / ref.abs().max())
    print(json.dumps({"blocks": name, "per_block_loop_with_wait_tflops": round(flops / t0 / 1e12, 3), "grouped_tflops": round(flops / t1 / 1e12, 3),
                      "speedup": round(t0 / t1, 2), "dispatches": groups, "rel_err_vs_torch": err}), flush=True)
