"""Executed warp-instructions per op of a run-time specialised kernel: walks the SASS of an ncu report in address order and
files every instruction under the last line of the generated kernel source (opm_jit_trace.cu) seen so far.
usage: ncu -i X.ncu-rep --page source --csv > sass.csv; ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > lines.csv
       python profiles/ncu_ops.py sass.csv lines.csv generated_kernel.cu [warps]"""
import collections
import csv
import sys
sass, lines, gen = sys.argv[1:4]
warps = float(sys.argv[4]) if len(sys.argv) > 4 else 312500.0
rows = list(csv.reader(open(lines)))
addr2line, cur, line = {}, None, None
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if len(r) > 10:
        if r[0] not in ('', 'Line No'):
            line = (cur, int(r[0]))
        elif r[0] == '' and r[2].startswith('0x'):
            addr2line.setdefault(int(r[2], 16), []).append(line)
rows = list(csv.reader(open(sass)))
h = next(r for r in rows if "Instructions Executed" in r)
ia, isrc, ie = h.index('Address'), h.index('Source'), h.index('Instructions Executed')
data = [r for r in rows if len(r) == len(h) and r != h]
src = open(gen).read().splitlines()
seg, per, perfp = None, collections.OrderedDict(), collections.Counter()
for r in data:
    a, n = int(r[ia], 16), int(r[ie])
    for l in addr2line.get(a, []):
        if l and l[0] == 'opm_jit_trace.cu':
            seg = l[1]
    per[seg] = per.get(seg, 0) + n
    op = r[isrc].split()
    op = (op[1] if op[0].startswith('@') else op[0]).split('.')[0]
    if op in ('DFMA', 'DMUL', 'DADD', 'DSETP', 'MUFU'):
        perfp[seg] += n
tot = sum(per.values())
print(f"total {tot / warps:.1f} warp-instructions per warp of rays")
for k, v in per.items():
    if v / tot > 0.003:
        print(f"{v / warps:8.1f} fp64 {perfp[k] / warps:7.1f}  L{k}: {src[k - 1].strip()[:120] if k else ''}")
