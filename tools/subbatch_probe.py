"""Is a sub-batched pipeline less efficient on the device than one big batch?  (device-resident reads, one context)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import minimap2_rs_b200 as mm2
from tools import gen
g = gen.genome(0xB2000002, 145_138_636)
goffs = np.array([0, g.size], dtype=np.uint64)
N = 100_000
cat, roffs = gen.reads(0xB2001002, g, goffs, N, 10_000, 0.0333, 0.0333, 0.0333)
ctx = mm2.Context(0)
gi = mm2.Index.build(ctx, g, goffs, ["chr8"])
d_cat = torch.empty(cat.size + 64, dtype=torch.uint8, device="cuda"); d_cat[:cat.size].copy_(torch.from_numpy(cat))
def run(nsub):
    per = N // nsub
    offs_l, d_offs = [], []
    for s in range(nsub):
        o = (roffs[s * per:(s + 1) * per + 1] - roffs[s * per]).astype(np.uint64)
        offs_l.append(o); d_offs.append(torch.from_numpy(o.astype(np.int64)).cuda())
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        tm = {}
        for s in range(nsub):
            r = ctx.map_batch(gi, None, offs_l[s], device_ptrs=(d_cat.data_ptr() + int(roffs[s * per]), d_offs[s].data_ptr()))
            for k, v in ctx.last_timings().items(): tm[k] = tm.get(k, 0) + v
            r.close()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print("nsub=%2d  wall %.1f ms  stage sum %.1f ms  %s" % (nsub, best * 1e3, sum(tm.values()), {k: round(v, 1) for k, v in tm.items()}), flush=True)
for nsub in (1, 2, 4, 8, 16, 32):
    run(nsub)
