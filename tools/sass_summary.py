#!/usr/bin/env python
"""Per-kernel SASS evidence from the built library: counts of the mnemonics that prove which hardware path a kernel uses
(UTMALDG/UTMASTG = TMA, UTC*MMA + LDTM/STTM = tcgen05 + TMEM, DMMA/HMMA = mma.sync tensor pipe, LDGSTS = cp.async, SYNCS = mbarrier).
    python tools/sass_summary.py [lib.so] > profiles/r2/sass_summary.txt
Runs cuobjdump -sass (CPU only, no GPU needed)."""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "einsums_b200", "lib", "libeinsums_b200.so")
KEYS = ["UTMALDG", "UTMASTG", "UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "LDTM", "STTM", "DMMA", "HMMA", "LDGSTS", "SYNCS", "LDS", "STS", "LDG", "STG", "DFMA", "FFMA", "FFMA2", "BAR"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for k in KEYS:
            if op == k or (k in ("UTMALDG", "UTMASTG", "UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM") and op.startswith(k)):
                counts[cur][k] += 1
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}: instruction counts per kernel (static, not executed counts)")
print("# " + " ".join(f"{k:>7s}" for k in KEYS) + "   total  kernel")
for fn, c in counts.items():
    name = demangle(fn)
    name = re.sub(r"\(anonymous namespace\)::|eb200::|unnamed>::", "", name)
    name = re.sub(r"\(.*$", "", name)
    print("  " + " ".join(f"{c[k]:7d}" for k in KEYS) + f" {total[fn]:7d}  {name[:150]}")
