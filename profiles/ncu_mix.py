"""Summarise an `ncu --page source --csv` export: SASS opcode mix weighted by executed count, and stall reasons.
usage: ncu -i X.ncu-rep --page source --csv > src.csv ; python profiles/ncu_mix.py src.csv [top-n]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(r for r in rows if "Instructions Executed" in r)
data = [r for r in rows if len(r) == len(h) and r != h]
ia, ie, isamp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
mix, samp, tot = collections.Counter(), collections.Counter(), 0
for r in data:
    toks = r[ia].split()
    if not toks:
        continue
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    n = int(r[ie])
    mix[op] += n
    tot += n
    samp[op] += int(r[isamp])
print(f"SASS instructions: {len(data)}   warp-instructions executed: {tot}")
for op, n in mix.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f"{op:12s} {n / tot:6.3f} {n:14d}   samples {samp[op]}")
st = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
tots = {c: sum(int(r[h.index(c)]) for r in data) for c in st}
s = sum(tots.values()) or 1
print("stall reasons (all samples):")
for c, v in sorted(tots.items(), key=lambda x: -x[1])[:10]:
    print(f"  {c:26s} {v:8d} {v / s:6.3f}")
