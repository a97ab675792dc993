#!/usr/bin/env python
"""Opcode mix and hottest instructions of one kernel from `ncu --page source --csv`."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0].startswith("0x")]
ie = idx['Instructions Executed']
tot = sum(int(r[ie]) for r in data)
print("SASS instructions:", len(data), "executed warp-inst:", tot)
ops = collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[idx['Source']])
    ops[m.group(2).split('.')[0] if m else '?'] += int(r[ie])
for op, c in ops.most_common(22):
    print(f"  {op:10s} {c:12d} {100*c/tot:5.1f}%")
print("hottest by stall samples:")
for s in sorted(((int(r[idx['# Samples']]), r[idx['Source']].strip()) for r in data), reverse=True)[:12]:
    print("  ", s)
