#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: share of samples / instructions per
48-instruction region of the LARGEST launch in the file, with the memory ops of the region."""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) >= 9 and r[0].startswith('0x')]
launches, cur, seen = [], [], set()
for r in rows:
    if r[0] in seen:
        launches.append(cur); cur, seen = [], set()
    seen.add(r[0]); cur.append(r)
launches.append(cur)
rows = max(launches, key=lambda l: sum(int(r[5]) for r in l))
base = int(rows[0][0], 16)
tot_s = sum(int(r[4]) for r in rows); tot_i = sum(int(r[5]) for r in rows)
print("launches", len(launches), "samples", tot_s, "warp-inst", tot_i, "sass", len(rows))
step = int(sys.argv[2]) if len(sys.argv) > 2 else 48
for i in range(0, len(rows), step):
    ch = rows[i:i + step]
    s = sum(int(r[4]) for r in ch); n = sum(int(r[5]) for r in ch)
    ops = [(r[1].split()[1] if r[1].strip().startswith('@') else r[1].split()[0]) for r in ch]
    key = [o for o in ops if o.startswith(('ATOMS', 'BAR', 'CALL', 'LDG', 'STG', 'RET', 'STS', 'ATOMG', 'RED', 'EXIT', 'MATCH'))]
    thr = sum(float(r[8]) * int(r[5]) for r in ch) / max(1, n)
    print(f"{int(ch[0][0],16)-base:6x} samples {100*s/max(1,tot_s):5.1f}% inst {100*n/max(1,tot_i):5.1f}% thr {thr:5.1f}  {' '.join(key[:9])}")
