"""Small dense-anchor workload (configs[4]-shaped: repeat-rich genome, 100 kb reads at 8 % error) for profiling the CTA-per-read
chaining kernel.  With MM2_LIB_PATH=minimap2_rs_b200/libmm2b200_prof.so (make -C minimap2_rs_b200/csrc prof) the library prints
the cycle counters of the dense pipeline on stderr."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minimap2_rs_b200 as mm2
from tools import workloads as wl
nreads = int(sys.argv[1]) if len(sys.argv) > 1 else 148
rlen = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
glen = int(float(sys.argv[3]) * 1e6) if len(sys.argv) > 3 else wl.C1_LEN
g, offs, names = wl.genome_c5(glen)
ctx = mm2.Context(0)
gi = mm2.Index.build(ctx, g, offs, names)
rc, ro = wl.reads("c5", g, offs, nreads, rlen)
for it in range(2):
    ctx.count_cells(it == 1)
    t0 = time.perf_counter()
    res = ctx.map_batch(gi, rc, ro)
    dt = time.perf_counter() - t0
    tm = ctx.last_timings()
    print("reads", nreads, "anchors", res.stats["n_anchors"], "rescued", res.stats["n_rescued"], "wall %.2f s" % dt, {k: round(v, 1) for k, v in tm.items()},
          ("cells %d = %.1f per anchor, %.1f Gcells/s" % (ctx.last_cells, ctx.last_cells / max(1, res.stats["n_anchors"]), ctx.last_cells / 1e9 / (tm.get("chain", 1) / 1e3))) if it else "", flush=True)
    res.close()
