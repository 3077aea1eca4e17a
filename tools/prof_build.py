"""Stage timings of one index build (single GPU) at a chosen genome size: python tools/prof_build.py [--config c3|c2] [--k 15]"""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minimap2_rs_b200 as mm2
from tools import workloads as wl
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c3")
ap.add_argument("--k", type=int, default=15)
ap.add_argument("--emulated", type=int, default=0, help="also run the sharded build's phases for this many virtual ranks")
a = ap.parse_args()
pins = []
def pinned(n):
    pb = mm2.PinnedBuffer(n); pins.append(pb); return pb.array(np.uint8, n)
if a.config == "c3":
    g, offs, names = wl.genome_c3(alloc=pinned)
else:
    g0, offs, names = wl.genome_c2(); g = pinned(g0.size); g[:] = g0
ctx = mm2.Context(0)
for it in range(3):
    t0 = time.perf_counter()
    gi = mm2.Index.build(ctx, g, offs, names, w=10, k=a.k)
    dt = time.perf_counter() - t0
    print("build %d: wall %.1f ms" % (it, dt * 1e3), {k_: round(v, 2) for k_, v in ctx.last_timings().items()}, gi.stats()[0], flush=True)
    gi.close()
if a.emulated:
    for it in range(2):
        t0 = time.perf_counter()
        gi = mm2.Index.build_sharded_emulated(ctx, a.emulated, g, offs, names, w=10, k=a.k)
        print("emulated x%d: wall %.1f ms" % (a.emulated, (time.perf_counter() - t0) * 1e3), {k_: round(v, 2) for k_, v in ctx.last_timings().items()}, flush=True)
        gi.close()
