"""SASS size per source line of one kernel (needs -lineinfo), optionally weighted by the executed counts of an
`ncu --page source --csv` export of the same build (joined by instruction order).
usage: python profiles/sass_lines.py <object-or-so> <kernel-substring> [top-n] [ncu-source.csv]"""
import collections
import csv
import re
import subprocess
import sys
import tempfile
from pathlib import Path

obj, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ncu_csv = sys.argv[4] if len(sys.argv) > 4 else None
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", str(Path(obj).resolve())], cwd=d, check=True, stdout=subprocess.DEVNULL)
    text = ""
    for cubin in Path(d).glob("*.cubin"):
        text += subprocess.run(["nvdisasm", "--print-line-info", str(cubin)], capture_output=True, text=True).stdout
cur, func = None, None
insts = collections.defaultdict(list)  # function -> [(line key, opcode)]
for ln in text.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        func = m.group(1)
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and func and pat in func:
        insts[func].append((cur, m.group(1)))
weights = None
if ncu_csv:
    rows = list(csv.reader(open(ncu_csv)))
    h = next(r for r in rows if "Instructions Executed" in r)
    ie = h.index("Instructions Executed")
    # the export concatenates launches: take the first launch's block
    first = []
    for r in rows[rows.index(h) + 1:]:
        if len(r) != len(h) or r == h:
            break
        first.append(int(r[ie]))
    weights = first
for f, lst in insts.items():
    print(f"{len(lst):7d} instructions ({len(lst) * 16 / 1024:.0f} KB)  {f[-70:]}")
    cnt, ops = collections.Counter(), collections.Counter()
    use_w = weights is not None and len(weights) == len(lst)
    if weights is not None and not use_w:
        print(f"   (ncu export has {len(weights)} instructions, this function {len(lst)}: static counts only)")
    for k, (key, op) in enumerate(lst):
        w = weights[k] if use_w else 1
        cnt[key] += w
        ops[op] += w
    tot = sum(cnt.values()) or 1
    print("   weighted by executed warp-instructions" if use_w else "   static instruction counts")
    print("   opcodes:", [(o, round(n / tot, 3)) for o, n in ops.most_common(12)])
    for key, n in cnt.most_common(top):
        print(f"   {n / tot:6.3f} {n:12d}  {key}")
