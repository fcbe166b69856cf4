#!/usr/bin/env python3
"""Text view of a ROFL_TIMELINE dump: per stream, the kernels in start order with start / duration; plus busy-union statistics."""
import sys, collections
rows = []
for l in open(sys.argv[1]):
    if l.startswith("#") or not l.strip(): continue
    f = l.split(); n, s, b, t0, t1 = f[:5]; rows.append((n, s, int(b), float(t0), float(t1), float(f[5]) if len(f) > 5 else float("nan")))
streams = collections.OrderedDict()
for r in sorted(rows, key=lambda r: r[3]): streams.setdefault(r[1], []).append(r)
end = max(r[4] for r in rows); start = min(r[3] for r in rows)
print("span %.2f ms, %d launches, %d streams" % (end - start, len(rows), len(streams)))
for i, (s, rs) in enumerate(streams.items()):
    busy = sum(r[4] - r[3] for r in rs)
    print("stream %d %s: %d launches, busy %.2f ms, first %.2f last %.2f" % (i, s, len(rs), busy, rs[0][3], rs[-1][4]))
if len(sys.argv) > 2:
    sid = {s: i for i, s in enumerate(streams)}
    for r in sorted(rows, key=lambda r: r[3]):
        print("%8.3f %8.3f  s%-2d %-22s blocks %-6d dur %.3f  enq %.3f" % (r[3], r[4], sid[r[1]], r[0], r[2], r[4] - r[3], r[5]))
