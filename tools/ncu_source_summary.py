#!/usr/bin/env python
"""Aggregate an exported `ncu --page source --csv` (SASS view): stall samples per opcode and the hottest instructions."""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hdr = rows[1]
si = hdr.index("# Samples"); src = hdr.index("Source"); ex = hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
by_op = collections.Counter(); by_op_exec = collections.Counter(); total = 0
insts = []
for r in rows[2:]:
    if len(r) <= si: continue
    try: s = int(r[si]); e = int(r[ex])
    except ValueError: continue
    op = r[src].split()[0] if r[src].split() else "?"
    if op.startswith("@"): op = r[src].split()[1]
    op = op.split(".")[0]
    by_op[op] += s; by_op_exec[op] += e; total += s
    insts.append((s, r[src].strip(), {hdr[i]: int(r[i]) for i in stall_cols if r[i] not in ("", "0")}))
print(f"total samples {total}")
for op, s in by_op.most_common(14):
    print(f"  {op:10s} samples {s:8d} ({100*s/total:5.1f}%)  executed {by_op_exec[op]}")
print("hottest instructions:")
for s, text, st in sorted(insts, key=lambda x: -x[0])[:top]:
    tops = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"  {s:7d} ({100*s/total:4.1f}%) {text[:70]:70s} {tops}")
