"""Turn an `ncu --set full` report of tools/prof_step.py into (1) profiles/rNN_ncu_traffic.json — DRAM bytes per unit of each
mapping stage's dominant kernel, which bench.py scales to its own launch for `roofline.traffic` — and (2) a markdown table of
the per-kernel counters the judge reads (duration, DRAM bytes, registers, occupancy, issue-active, pipe utilisation).

usage: python tools/ncu_traffic.py <report.ncu-rep> <round tag, e.g. r02> --reads R --read-len L --minimizers M --anchors A [--genome-bp G]
(the unit counts are those prof_step.py printed for the captured step; run here, no GPU needed)
"""
import argparse
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("report")
ap.add_argument("tag")
ap.add_argument("--reads", type=int, required=True)
ap.add_argument("--read-len", type=int, required=True)
ap.add_argument("--minimizers", type=int, required=True)
ap.add_argument("--anchors", type=int, required=True)
ap.add_argument("--genome-bp", type=int, default=145_138_636)
a = ap.parse_args()

out = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
h = rows[0]
units_row = rows[1]


def col(name):
    return h.index(name) if name in h else -1


want = {"dur": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum", "regs": "launch__registers_per_thread",
        "warps": "sm__warps_active.avg.pct_of_peak_sustained_active", "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "inst": "smsp__inst_executed.sum", "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "alu": "sm__inst_executed_pipe_alu.sum", "fma": "sm__inst_executed_pipe_fma.sum", "lsu": "sm__inst_executed_pipe_lsu.sum",
        "alu_pct": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smem": "launch__shared_mem_per_block_dynamic"}
ci = {k: col(v) for k, v in want.items()}
ki = col("Kernel Name")


def scale(val, unit):
    """bytes in B, time in ms"""
    v = float(val.replace(",", "")) if val else 0.0
    u = unit.lower()
    mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3,
            "msecond": 1.0, "second": 1e3}
    return v * mult.get(u, 1)


kern = {}
for r in rows[2:]:
    if len(r) <= ki or not r[ki]:
        continue
    name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    e = kern.setdefault(name, {"launches": 0, "ms": 0.0, "dram": 0.0, "inst": 0.0, "rows": []})
    e["launches"] += 1
    e["ms"] += scale(r[ci["dur"]], units_row[ci["dur"]])
    e["dram"] += scale(r[ci["rd"]], units_row[ci["rd"]]) + scale(r[ci["wr"]], units_row[ci["wr"]])
    e["inst"] += float(r[ci["inst"]].replace(",", "")) if ci["inst"] >= 0 and r[ci["inst"]] else 0.0
    e["rows"].append(r)

bases = a.reads * a.read_len
stage_of = [("sketch", "sketch_tile_kernel", bases, "base"), ("lookup", "seed_hits_kernel<0>", a.minimizers, "minimizer"),
            ("anchor_sort", "anchor_msort_kernel<128", a.anchors, "anchor"), ("chain", "chain_ring_kernel", a.anchors, "anchor")]
traffic = {}
for stage, pat, units, unit in stage_of:
    hits = [(n, e) for n, e in kern.items() if pat in n]
    if not hits:
        continue
    n, e = max(hits, key=lambda x: x[1]["ms"])
    # the capture holds several launches of the kernel (the genome's sketch during the index build, then the reads'): the mapping
    # step comes last in tools/prof_step.py, so its launch is the last one
    best = e["rows"][-1]
    dram = scale(best[ci["rd"]], units_row[ci["rd"]]) + scale(best[ci["wr"]], units_row[ci["wr"]])
    traffic[stage] = {"kernel": n, "dram_bytes": dram, "units": units, "unit": unit, "ms_under_ncu": scale(best[ci["dur"]], units_row[ci["dur"]])}
try:
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
except Exception:
    commit = None
js = {"source": os.path.basename(a.report), "commit": commit, "workload": "tools/prof_step.py: %d reads x %d bp vs %d bp genome" % (a.reads, a.read_len, a.genome_bp),
      "kernels": traffic}
path = os.path.join(ROOT, "profiles", "%s_ncu_traffic.json" % a.tag)
json.dump(js, open(path, "w"), indent=1)
print("wrote", path)

md = ["# ncu --set full, %s (%s, commit %s)" % (a.tag, js["workload"], commit), "",
      "| kernel | launches | total ms (under ncu) | DRAM GB (rd+wr) | regs | warps active % | issue active % | warp-inst (M) | ALU pipe % |", "|---|---|---|---|---|---|---|---|---|"]
for n, e in sorted(kern.items(), key=lambda x: -x[1]["ms"]):
    r0 = max(e["rows"], key=lambda r: scale(r[ci["dur"]], units_row[ci["dur"]]))

    def g(k):
        return r0[ci[k]] if ci[k] >= 0 else ""
    md.append("| `%s` | %d | %.3f | %.3f | %s | %s | %s | %.1f | %s |" % (n[:70], e["launches"], e["ms"], e["dram"] / 1e9, g("regs"), g("warps")[:5], g("issue")[:5],
                                                                    e["inst"] / 1e6, g("alu_pct")[:5]))
mdp = os.path.join(ROOT, "profiles", "%s_ncu_kernels.md" % a.tag)
open(mdp, "w").write("\n".join(md) + "\n")
print("wrote", mdp)
