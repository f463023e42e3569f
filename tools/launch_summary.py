#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean us, share."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[hdr]; ki, vi = H.index("Kernel Name"), H.index("Metric Value")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
seq = [(r[ki].split("(")[0].replace("(anonymous namespace)::", "").replace("<unnamed>::", ""), float(r[vi].replace(",", "")) / 1e3)
       for r in rows[hdr + 1:] if len(r) > vi][skip:]
tot = sum(v for _, v in seq)
agg = collections.OrderedDict()
for k, v in seq:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
print(f"{'kernel':28s} {'launches':>8s} {'mean_us':>10s} {'total_us':>10s} {'share':>7s}")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:28s} {c:8d} {v / c:10.1f} {v:10.1f} {100 * v / tot:6.1f}%")
print(f"{'TOTAL':28s} {len(seq):8d} {'':10s} {tot:10.1f}")
if len(sys.argv) > 3:
    print("\nlast pipeline, in order:")
    for k, v in seq[-int(sys.argv[3]):]: print(f"  {k:28s} {v:10.1f} us")
