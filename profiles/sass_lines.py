"""Static SASS size per source line of one kernel (needs -lineinfo).
usage: python profiles/sass_lines.py <object-or-so> <kernel-substring> [top-n]
Extracts the cubin with cuobjdump, disassembles with `nvdisasm --print-line-info` and counts instructions per (file, line)."""
import collections
import re
import subprocess
import sys
import tempfile
from pathlib import Path

obj, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", str(Path(obj).resolve())], cwd=d, check=True, stdout=subprocess.DEVNULL)
    text = ""
    for cubin in Path(d).glob("*.cubin"):
        text += subprocess.run(["nvdisasm", "--print-line-info", str(cubin)], capture_output=True, text=True).stdout
cur, func = None, None
cnt, per_func, ops = collections.Counter(), collections.Counter(), collections.Counter()
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
        cnt[cur] += 1
        per_func[func[-70:]] += 1
        ops[m.group(1)] += 1
for f, n in per_func.items():
    print(f"{n:7d} instructions ({n * 16 / 1024:.0f} KB)  {f}")
by_file = collections.Counter()
for (f, _), n in ((k, v) for k, v in cnt.items() if k):
    by_file[f] += n
print("by file:", by_file.most_common(8))
print("opcodes:", ops.most_common(14))
for k, n in cnt.most_common(top):
    print(f"{n:6d}  {k}")
