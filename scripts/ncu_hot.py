"""Summarise an `ncu --page source --csv` dump: top instructions by stall samples and totals per
stall reason / per opcode."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
S = col["# Samples"]
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[S] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
byr = collections.Counter()
byop = collections.Counter()
for r in data:
    n = int(r[S] or 0)
    op = r[col["Source"]].split()[0] if r[col["Source"]].split() else "?"
    if op.startswith("@"):
        op = r[col["Source"]].split()[1]
    byop[op.split(".")[0]] += n
    for h in reasons:
        byr[h] += int(r[col[h]] or 0)
print("by reason:", [(k, v) for k, v in byr.most_common(8)])
print("by opcode:", byop.most_common(14))
top = sorted(data, key=lambda r: -int(r[S] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for r in top:
    rs = sorted(((int(r[col[h]] or 0), h) for h in reasons), reverse=True)[:2]
    print(f"{int(r[S]):6d} {100 * int(r[S]) / tot:5.1f}%  exec {r[col['Instructions Executed']]:>9}  {r[col['Source']].strip()[:70]:70s} {rs}")
