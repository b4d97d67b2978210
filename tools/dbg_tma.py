import os, sys
sys.path.insert(0, '/root/repo')
import torch
from einsums_b200 import ops
ops.init(0)
torch.manual_seed(0)
o, v = 20, int(os.environ.get("V", 80))
g = torch.rand(o, o, v, v, dtype=torch.float64, device="cuda") * 2 - 1
t = torch.rand(o, o, v, v, dtype=torch.float64, device="cuda") * 2 - 1
W0 = torch.rand(v, v, v, v, dtype=torch.float64, device="cuda") * 2 - 1
ref = W0 + 0.25 * (t.reshape(o * o, -1).T @ g.reshape(o * o, -1)).reshape(v, v, v, v)
for rep in range(int(os.environ.get('REPS', 10))):
    W = W0.clone()
    ops.contract("abef", W, "mnef", g, "mnab", t, 0.25, 1.0)
    torch.cuda.synchronize()
    diff = (W - ref).abs().reshape(v * v, v * v)
    bad = (diff > 1e-9)
    print("rep", rep, "max diff", diff.max().item(), "bad count", bad.sum().item(), "of", diff.numel())
    if bad.any():
        idx = bad.nonzero()
        rows, cols = idx[:, 0], idx[:, 1]   # rows = (a,b) = N index, cols = (e,f) = M index
        print("  bad N-tiles (64):", torch.unique(rows // 64)[:20].tolist(), " bad M-tiles (128):", torch.unique(cols // 128)[:20].tolist())
        print("  rows in tile:", torch.unique(rows % 64)[:40].tolist())
        print("  cols in tile:", torch.unique(cols % 128)[:40].tolist())
        print("  sample diffs:", diff[bad][:5].tolist())
        tiles = torch.unique(torch.stack([rows // 64, cols // 128], 1), dim=0)
        for tn, tm in tiles.tolist():
            sel = (rows // 64 == tn) & (cols // 128 == tm)
            rr, cc = rows[sel] % 64, cols[sel] % 128
            print(f"   tile n={tn} m={tm}: count {sel.sum().item()} rows {rr.min().item()}..{rr.max().item()} cols {cc.min().item()}..{cc.max().item()}")
