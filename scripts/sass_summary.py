"""Per-kernel counts of the Blackwell-native SASS mnemonics in the built library (no GPU needed):

    python scripts/sass_summary.py > profiles/sass_summary.txt

UTCHMMA / UTCQMMA = tcgen05.mma (kind::f16 / other kinds), UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st
(TMEM), UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk (1-D bulk copy), UTMAPF = TMA L2
prefetch, SYNCS = mbarrier ops, HMMA = legacy mma.sync (the decode strip GEMM only), REDG/RED = global reductions."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "musicgeneration_b200", "libmt_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "HMMA", "RED", "MUFU.EX2", "FFMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k + ".") or (k == "RED" and op.startswith("REDG")):
                counts[cur][k] += 1
        counts[cur]["_n"] += 1
    names = list(counts)
    try:
        dm = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        demangle = dict(zip(names, dm))
    except Exception:
        pass
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} -- per-kernel instruction counts (static), sm_100a")
    print("# " + " ".join(f"{k:>8}" for k in ["instrs"] + KEYS) + "  kernel")
    for n in names:
        c = counts[n]
        short = demangle.get(n, n).replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("(int)", "").replace("(bool)", "")
        short = re.sub(r"\(.*", "", short).replace("mt::", "").replace("void ", "")
        print("  " + " ".join(f"{c[k]:>8}" for k in ["_n"] + KEYS) + "  " + short[:110])
    tc = [n for n in names if counts[n]["UTCHMMA"] or counts[n]["UTCQMMA"]]
    print(f"# {len(tc)} of {len(names)} kernels issue tcgen05.mma; "
          f"{sum(1 for n in names if counts[n]['UTMALDG'])} use TMA tensor loads; "
          f"{sum(1 for n in names if counts[n]['LDTM'])} read TMEM")


if __name__ == "__main__":
    sys.exit(main())
