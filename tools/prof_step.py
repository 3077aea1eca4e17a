"""Small fixed workload for ncu: one index build + mapping steps (used for the launch list and --set full captures)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minimap2_rs_b200 as mm2
from tools import gen

ap = argparse.ArgumentParser()
ap.add_argument("--genome-mbp", type=float, default=145.138636)
ap.add_argument("--reads", type=int, default=20000)
ap.add_argument("--read-len", type=int, default=10000)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--builds", type=int, default=2)
a = ap.parse_args()
glen = int(a.genome_mbp * 1e6)
g = gen.genome(0xB2000002, glen)
goffs = np.array([0, glen], dtype=np.uint64)
cat, roffs = gen.reads(0xB2001002, g, goffs, a.reads, a.read_len, 0.0333, 0.0333, 0.0333)
ctx = mm2.Context(0)
for _ in range(a.builds):
    gi = mm2.Index.build(ctx, g, goffs, ["chr8"])
    print("build", gi.build_timings())
for _ in range(a.steps):
    res = ctx.map_batch(gi, cat, roffs)
    print("map", {k: round(v, 3) for k, v in ctx.last_timings().items()}, res.stats)
    res.close()
