"""e2e timing of mm2_map_batch / mm2_map_batch_packed (pinned host buffers) for different sub-batch sizes / worker counts
(the environment is read per Context)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minimap2_rs_b200 as mm2
from tools import gen
g = gen.genome(0xB2000002, 145_138_636)
goffs = np.array([0, g.size], dtype=np.uint64)
cat, roffs = gen.reads(0xB2001002, g, goffs, 100_000, 10_000, 0.0333, 0.0333, 0.0333)
pin = mm2.PinnedBuffer(cat.size); pr = pin.array(np.uint8, cat.size); pr[:] = cat
pin2 = mm2.PinnedBuffer(cat.size // 4 + 128); pk = pin2.array(np.uint8, cat.size // 4 + 128)
_, n_pos = mm2.pack_reads(pr, out=pk)
c0 = mm2.Context(0)
gi = mm2.Index.build(c0, g, goffs, ["chr8"])
for workers, mb in [(4, 64), (4, 32), (4, 128), (4, 256), (3, 128), (2, 256)]:
    os.environ["MM2_WORKERS"] = str(workers); os.environ["MM2_SUBBATCH_MB"] = str(mb); os.environ["MM2_PIPELINE"] = "1"
    c = mm2.Context(0)
    out = []
    for fn in (lambda: c.map_batch(gi, pr, roffs), lambda: c.map_batch_packed(gi, pk, n_pos, roffs)):
        for _ in range(2): fn().close()
        t0 = time.perf_counter()
        for _ in range(4): fn().close()
        out.append((time.perf_counter() - t0) / 4)
    print("workers=%d subbatch=%dMB  ascii %.1f ms  packed %.1f ms" % (workers, mb, out[0] * 1e3, out[1] * 1e3), flush=True)
    c.close()
