"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = r[kn].split("(")[0][:70]
    v = float(r[mv].replace(",", ""))
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v for _, v in agg.values())
unit = rows[hi + 1][hdr.index("Metric Unit")]
print(f"{'kernel':72s} {'n':>5s} {'total ' + unit:>14s} {'share':>7s} {'avg':>10s}")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:72s} {n:5d} {v:14.0f} {v / tot * 100:6.1f}% {v / n:10.0f}")
print(f"{'TOTAL':72s} {sum(n for n, _ in agg.values()):5d} {tot:14.0f}")
