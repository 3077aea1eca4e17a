"""Small dense-anchor workload (repeat-rich genome, long reads) for profiling the chain kernel on configs[4]-like input."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minimap2_rs_b200 as mm2
from tools import gen
g = gen.repeat_genome(0xB2000005, 10_000_000, 0.4, 0.2)
offs = np.array([0, g.size], dtype=np.uint64)
ctx = mm2.Context(0)
gi = mm2.Index.build(ctx, g, offs, ["rep"])
nreads = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rlen = int(sys.argv[2]) if len(sys.argv) > 2 else 30000
rc, ro = gen.reads(0xB2001005, g, offs, nreads, rlen, 0.027, 0.027, 0.026)
for _ in range(2):
    t0 = time.perf_counter()
    res = ctx.map_batch(gi, rc, ro)
    print("reads", nreads, "anchors", res.stats["n_anchors"], "wall %.2f s" % (time.perf_counter() - t0), {k: round(v, 1) for k, v in ctx.last_timings().items()}, flush=True)
    res.close()
