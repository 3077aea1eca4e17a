"""Static SASS instruction count per CUDA source line of one kernel (nvdisasm -g on an object file built with -lineinfo).
usage: python tools/sass_lines.py sketch.o 'v4IjLi10' [src.cu]   -> instructions per source line, in source order"""
import collections
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
src = sys.argv[3] if len(sys.argv) > 3 else None
import os
import tempfile
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
cur_fn, cur_line = None, None
cnt = collections.Counter()
ops = collections.defaultdict(collections.Counter)
for ln in out.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        cur_fn = m.group(1)
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if cur_fn and pat in cur_fn:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
        if m and cur_line:
            cnt[cur_line] += 1
            ops[cur_line][m.group(2).split(".")[0]] += 1
lines = open(src).read().splitlines() if src else None
tot = sum(cnt.values())
print("total", tot)
for (f, l), c in sorted(cnt.items()):
    text = lines[l - 1].strip()[:110] if lines and f == src.split("/")[-1] else ""
    print("%s:%d\t%d\t%s\t%s" % (f, l, c, " ".join("%s:%d" % kv for kv in ops[(f, l)].most_common(5)), text))
