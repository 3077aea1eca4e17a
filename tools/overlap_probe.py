"""Does an H2D copy overlap with our kernels?  thread A: device-resident map_batch loop; thread B: pinned H2D copies (torch)."""
import os, sys, time, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import minimap2_rs_b200 as mm2
from tools import gen
g = gen.genome(0xB2000002, 145_138_636)
goffs = np.array([0, g.size], dtype=np.uint64)
N = 50_000
cat, roffs = gen.reads(0xB2001002, g, goffs, N, 10_000, 0.0333, 0.0333, 0.0333)
ctx = mm2.Context(0)
gi = mm2.Index.build(ctx, g, goffs, ["chr8"])
d_cat = torch.empty(cat.size + 64, dtype=torch.uint8, device="cuda"); d_cat[:cat.size].copy_(torch.from_numpy(cat))
d_off = torch.from_numpy(roffs.astype(np.int64)).cuda()
src = torch.empty(512 << 20, dtype=torch.uint8).pin_memory()
dst = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
s2 = torch.cuda.Stream()
def copy_once():
    with torch.cuda.stream(s2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s2); dst.copy_(src, non_blocking=True); e1.record(s2); e1.synchronize()
        return e0.elapsed_time(e1)
def map_once():
    t0 = time.perf_counter(); ctx.map_batch(gi, None, roffs, device_ptrs=(d_cat.data_ptr(), d_off.data_ptr())).close(); return (time.perf_counter() - t0) * 1e3
map_once(); copy_once()
map_once(); print("   alone:", {k: round(v, 2) for k, v in ctx.last_timings().items()})
print("alone: copy 512MB %.1f ms (%.1f GB/s); map 0.5 Gbase %.1f ms" % (copy_once(), 0.512 / copy_once() * 1e3 * 1.048, map_once()))
stop = False; copies = []
def copier():
    while not stop: copies.append(copy_once())
th = threading.Thread(target=copier); th.start()
maps = []
for _ in range(6):
    maps.append(map_once()); print("   under copy:", {k: round(v, 2) for k, v in ctx.last_timings().items()})
stop = True; th.join()
print("concurrent: map %.1f ms avg; copy %.1f ms avg over %d copies" % (np.mean(maps), np.mean(copies), len(copies)))
