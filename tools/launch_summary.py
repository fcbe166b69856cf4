#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and (optionally) the sequence."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
d = collections.defaultdict(lambda: [0, 0.0]); seq = []
for r in rows[hdr + 1:]:
    if len(r) < 15: continue
    name = r[4].split('(')[0][:40]
    d[name][0] += 1; d[name][1] += float(r[-1]) / 1e6; seq.append((name, r[7], r[8], float(r[-1]) / 1e6))
tot = sum(v[1] for v in d.values())
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for k, v in sorted(d.items(), key=lambda x: -x[1][1]): print(f"| {k} | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f}% |")
print(f"\ntotal {tot:.1f} ms")
if len(sys.argv) > 2:
    a, b = (int(x) for x in sys.argv[2].split(':'))
    for s in seq[a:b]: print(s)
