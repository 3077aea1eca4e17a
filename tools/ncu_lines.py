"""Attribute ncu per-SASS-instruction counters to CUDA source lines.

usage: python tools/ncu_lines.py <report.ncu-rep> <kernel regex> <cubin name substring> [min_pct]
Joins `ncu --page source --csv` (SASS view: instructions executed, stall samples) with `nvdisasm -g` line markers by
instruction order inside the kernel.  Needs -lineinfo at compile time.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, kre, cub = sys.argv[1], sys.argv[2], sys.argv[3]
min_pct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.7
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "minimap2_rs_b200", "libmm2b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if cub in f][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
kname = rows[0][1]
hdr = rows[1]
ci, si, so_ = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
ti = hdr.index("Thread Instructions Executed")
# the csv repeats the whole table once per matching launch: split on the repeated header, take launch MM2_NCU_LAUNCH (default: last)
blocks, curb = [], []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        blocks.append(curb); curb = []
        continue
    if len(r) > ci and r[ci].isdigit():
        curb.append((r[so_].strip(), int(r[ci]), int(r[si]), int(r[ti])))
blocks.append(curb)
blocks = [b for b in blocks if b]
ins = blocks[int(os.environ.get("MM2_NCU_LAUNCH", "-1"))]
# find the kernel's function in the disassembly: match by the mangled core name
core = re.sub(r"[^A-Za-z0-9_]", "", kname.split("(")[0].split("::")[-1].split("<")[0])
tmpls = re.findall(r"<([^<>]*)>", kname.split("(")[0])
tmpl = tmpls[-1] if tmpls else None
start = None
cands = [i for i, l in enumerate(dis) if l.startswith("_Z") and core in l and l.rstrip().endswith(":")]
mint = re.search(r"<\(int\)(\d+)>\(", kname)
if mint and len(cands) > 1:
    cands = [i for i in cands if "%sILi%sE" % (core, mint.group(1)) in dis[i]] or cands
elif tmpl and len(cands) > 1:
    t = tmpl
    key = {"unsigned int": "Ij", "unsigned long": "Im", "unsigned long long": "Iy"}.get(t)
    mi = re.fullmatch(r"\(int\)(\d+)", t.strip())
    if mi:
        key = "ILi%sE" % mi.group(1)
    if key:
        cands = [i for i in cands if core + key in dis[i]] or cands
start = cands[0]
lines, cur = [], None
for l in dis[start + 1:]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append((cur, l.split("*/", 1)[1].strip().rstrip(";").strip()))
    if l.startswith("_Z") or l.startswith("//-----"):
        break
n = min(len(lines), len(ins))
agg = {}
tot_i = sum(x[1] for x in ins)
tot_s = sum(x[2] for x in ins)
for k in range(n):
    key = lines[k][0]
    a = agg.setdefault(key, [0, 0, 0])
    a[0] += ins[k][1]; a[1] += ins[k][2]; a[2] += ins[k][3]
src_cache = {}
print("kernel: %s\ninstructions (warp-level): %d, samples: %d, sass lines matched: %d/%d" % (kname[:90], tot_i, tot_s, n, len(ins)))
for key, a in sorted(agg.items(), key=lambda kv: (kv[0] or ("", 0))):
    pi, ps = 100.0 * a[0] / max(1, tot_i), 100.0 * a[1] / max(1, tot_s)
    if pi < min_pct and ps < min_pct:
        continue
    text = ""
    if key:
        f = [os.path.join(dp, key[0]) for dp, _, fs in os.walk(root) if key[0] in fs and "gpurun_out" not in dp]
        if f:
            src_cache.setdefault(f[0], open(f[0], errors="replace").read().split("\n"))
            text = src_cache[f[0]][key[1] - 1].strip()[:100]
    print("%-14s inst %5.1f%%  stall-samples %5.1f%%  thr/inst %4.1f | %s" % ("%s:%d" % key if key else "?", pi, ps, a[2] / max(1, a[0]), text))
