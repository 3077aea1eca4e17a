"""e2e timing of mm2_map_batch (pinned host buffers) for different sub-batch sizes / worker counts (env is read per Context)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minimap2_rs_b200 as mm2
from tools import gen
g = gen.genome(0xB2000002, 145_138_636)
goffs = np.array([0, g.size], dtype=np.uint64)
cat, roffs = gen.reads(0xB2001002, g, goffs, 100_000, 10_000, 0.0333, 0.0333, 0.0333)
pin = mm2.PinnedBuffer(cat.size); pr = pin.array(np.uint8, cat.size); pr[:] = cat
c0 = mm2.Context(0)
gi = mm2.Index.build(c0, g, goffs, ["chr8"])
for workers, mb in [(4, 64), (4, 32), (4, 48), (4, 24), (3, 32), (4, 96), (2, 64)]:
    if workers:
        os.environ["MM2_WORKERS"] = str(workers); os.environ["MM2_SUBBATCH_MB"] = str(mb); os.environ["MM2_PIPELINE"] = "1"
    else:
        os.environ["MM2_PIPELINE"] = "0"
    c = mm2.Context(0)
    for _ in range(2): c.map_batch(gi, pr, roffs).close()
    t0 = time.perf_counter()
    for _ in range(4): c.map_batch(gi, pr, roffs).close()
    dt = (time.perf_counter() - t0) / 4
    print("workers=%d subbatch=%dMB  e2e %.1f ms  %.2f Gbase/s" % (workers, mb, dt * 1e3, 1.0 / dt), flush=True)
    c.close()
