"""Timeline of one end-to-end mapping call (MM2_TRACE=1) plus the raw pinned H2D rate of the box."""
import os, sys, time
os.environ["MM2_TRACE"] = "1"
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import minimap2_rs_b200 as mm2
from tools import gen
g = gen.genome(0xB2000002, 145_138_636)
goffs = np.array([0, g.size], dtype=np.uint64)
cat, roffs = gen.reads(0xB2001002, g, goffs, 100_000, 10_000, 0.0333, 0.0333, 0.0333)
pin = mm2.PinnedBuffer(cat.size); pr = pin.array(np.uint8, cat.size); pr[:] = cat
# raw H2D rate from the same pinned buffer
d = torch.empty(cat.size, dtype=torch.uint8, device="cuda")
src = torch.from_numpy(pr)
for _ in range(2): d.copy_(src, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter(); d.copy_(src, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("raw H2D of %.2f GB: %.2f ms = %.1f GB/s (is_pinned=%s)" % (cat.size / 1e9, dt * 1e3, cat.size / dt / 1e9, src.is_pinned()), flush=True)
c = mm2.Context(0)
gi = mm2.Index.build(c, g, goffs, ["chr8"])
for _ in range(3): c.map_batch(gi, pr, roffs).close()
sys.stderr.write("[mm2 trace] ==== timed call ====\n"); sys.stderr.flush()
t0 = time.perf_counter(); c.map_batch(gi, pr, roffs).close(); dt = time.perf_counter() - t0
sys.stderr.write("[mm2 trace] ==== end: %.2f ms ====\n" % (dt * 1e3))
