#!/usr/bin/env python
"""Top stall-sample instructions and per-opcode shares from `ncu --page source --csv`."""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr, data = rows[hi], rows[hi + 1:]
ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in data if len(r) > ci["# Samples"]]
tot = sum(int(r[ci["# Samples"]] or 0) for r in data)
print("total samples", tot)
for r in sorted(data, key=lambda r: -int(r[ci["# Samples"]] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    n = int(r[ci["# Samples"]])
    rs = sorted(((k, int(r[ci[k]] or 0)) for k in hdr if k.startswith("stall_") and "Not Issued" not in k),
                key=lambda kv: -kv[1])[:3]
    print(f"{n:7d} {100*n/tot:5.1f}% {r[ci['Source']][:64]:64s} {rs}")
agg = Counter()
for r in data:
    src = r[ci['Source']].split()
    op = (src[1] if src and src[0].startswith('@') else (src[0] if src else '?'))
    agg[op] += int(r[ci["# Samples"]] or 0)
print([(k, round(100 * v / tot, 1)) for k, v in agg.most_common(16)])
