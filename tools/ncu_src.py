"""Attribute the per-SASS-instruction counters of ONE kernel instance in an ncu report to CUDA source lines.

usage: python tools/ncu_src.py <report.ncu-rep> <kernel regex> <mangled-name substring> <source file> [units] [min_pct]
`units` (e.g. the number of bases or anchors the launch processed) turns counts into thread-instructions per unit.
Joins `ncu --page source --csv` (SASS view) with `nvdisasm -g` line markers by instruction order (needs -lineinfo).
"""
import csv, os, re, subprocess, sys, tempfile
rep, kre, sub, srcf = sys.argv[1:5]
units = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0
min_pct = float(sys.argv[6]) if len(sys.argv) > 6 else 0.5
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
inst, cur = [], None
for r in csv.reader(out.split("\n")):
    if r and r[0] == "Kernel Name":
        cur = []; inst.append(cur); continue
    if cur is not None:
        cur.append(r)
blk = inst[-1]
hdr = blk[0]
ci, si, ti = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
ins = [(r[1].strip(), int(r[ci]), int(r[si]), int(r[ti])) for r in blk[1:] if len(r) > ci and r[ci].isdigit()]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "minimap2_rs_b200", "libmm2b200.so")], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
lines = None
for f in os.listdir(tmp):
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout.split("\n")
    cands = [i for i, l in enumerate(dis) if l.startswith("_Z") and sub in l and l.rstrip().endswith(":")]
    if not cands:
        continue
    lines, curl = [], None
    for l in dis[cands[0] + 1:]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            curl = (os.path.basename(m.group(1)), int(m.group(2))); continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            lines.append((curl, l.split("*/", 1)[1].strip()))
        if l.startswith("_Z") or l.startswith("//-----"):
            break
    break
n = min(len(lines), len(ins))
mism = sum(1 for k in range(n) if lines[k][1].split()[0].strip(";") != ins[k][0].split()[0] and not ins[k][0].startswith("@"))
tot = sum(x[1] for x in ins); tots = sum(x[2] for x in ins); tott = sum(x[3] for x in ins)
print("instances %d; warp-instr %d (%.2f per unit), thread-instr per unit %.1f; sass %d vs disasm %d, opcode mismatches %d" %
      (len(inst), tot, tot / units if units else 0, tott / units if units else 0, len(ins), len(lines), mism))
agg = {}
for k in range(n):
    a = agg.setdefault(lines[k][0], [0, 0, 0]); a[0] += ins[k][1]; a[1] += ins[k][2]; a[2] += ins[k][3]
src = open(srcf, errors="replace").read().split("\n")
base = os.path.basename(srcf)
for key, a in sorted(agg.items(), key=lambda kv: kv[0] or ("", 0)):
    if 100.0 * a[0] / tot < min_pct and 100.0 * a[1] / max(1, tots) < min_pct:
        continue
    text = src[key[1] - 1].strip()[:90] if key and key[0] == base else (key[0] if key else "?")
    print("%4d inst %5.1f%% stall %5.1f%% %s| %s" % (key[1] if key else 0, 100.0 * a[0] / tot, 100.0 * a[1] / max(1, tots),
                                                   ("%6.1f thr-inst/unit " % (a[2] / units)) if units else "", text))
