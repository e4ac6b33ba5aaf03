"""Instruction mix of an `ncu --page source --csv` dump: executed warp-instructions per opcode,
normalised by a divisor (e.g. warps x steps) given as argv[2]."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
byop = collections.Counter()
samp = collections.Counter()
for r in rows[2:]:
    src = r[col["Source"]].split()
    if not src:
        continue
    op = src[1] if src[0].startswith("@") else src[0]
    op = ".".join(op.split(".")[:2])
    byop[op] += int(r[col["Instructions Executed"]] or 0)
    samp[op] += int(r[col["# Samples"]] or 0)
tot = sum(byop.values())
print("total warp-instr", tot, "per unit", round(tot / div, 1))
for k, v in byop.most_common(40):
    print(f"{k:28s} {v / div:8.1f}  samples {samp[k]}")
