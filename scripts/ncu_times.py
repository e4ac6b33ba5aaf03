"""Per-kernel averages of an `ncu --csv --metrics ...` launch list (the last of N repeats of each kernel is what counts)."""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
i = [k for k, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[i]
agg = OrderedDict()
for r in rows[i + 1:]:
    d = dict(zip(hdr, r))
    key = d["Kernel Name"][:60]
    agg.setdefault(key, {}).setdefault(d["Metric Name"], []).append(float(d["Metric Value"].replace(",", "")))
for k, m in agg.items():
    if "at::" in k:
        continue
    print(k)
    for name, v in m.items():
        print(f"    {name:80s} n={len(v):3d} last={v[-1]:.4g} mean={sum(v) / len(v):.4g}")
