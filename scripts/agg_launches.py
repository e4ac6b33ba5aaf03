"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    name = row["Kernel Name"]
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    if unit in ("ns", "nsecond"):
        v /= 1e3
    elif unit in ("ms", "msecond"):
        v *= 1e3
    elif unit in ("s", "second"):
        v *= 1e6
    name = re.sub(r"\(.*", "", name)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{v[1]:10.1f} us {100 * v[1] / tot:5.1f}%  n={v[0]:4d}  avg {v[1] / v[0]:8.1f} us  {k[:100]}")
print("total us", round(tot, 1), "launches", sum(v[0] for v in agg.values()))
