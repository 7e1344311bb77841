"""Executed warp-instructions and stall samples per CUDA source line of an `ncu --page source --print-source cuda,sass --csv` export.
usage: python profiles/ncu_lines.py export.csv [top-n]"""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cur, out = None, []
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if len(r) > 10 and r[0] not in ('', 'Line No'):
        try:
            n, s = int(r[7]), int(r[6])
        except ValueError:
            continue
        out.append((n, s, cur, int(r[0])))
tot = sum(o[0] for o in out)
print("warp-instructions executed:", tot)
out.sort(reverse=True)
for n, s, f, l in out[:top]:
    print(f"{n / tot:6.3f} {n:11d} samples {s:6d} {f}:{l}")
