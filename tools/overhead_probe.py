"""Where the wall time of one device-resident mapping step goes on the host side (C call vs Python wrapper)."""
import os, sys, time, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import minimap2_rs_b200 as mm2
from tools import gen
glen = 145138636
g = gen.genome(0xB2000002, glen)
goffs = np.array([0, glen], dtype=np.uint64)
cat, roffs = gen.reads(0xB2001002, g, goffs, 100000, 10000, 0.0333, 0.0333, 0.0333)
ctx = mm2.Context(0, stream=torch.cuda.current_stream().cuda_stream)
gi = mm2.Index.build(ctx, g, goffs, ["chr8"])
d_cat = torch.empty(cat.size + 64, dtype=torch.uint8, device="cuda"); d_cat[:cat.size].copy_(torch.from_numpy(cat))
d_off = torch.from_numpy(roffs.astype(np.int64)).cuda()
opts = mm2.default_map_opts(10, 15)
torch.cuda.synchronize()
L = mm2.lib()
for it in range(6):
    t0 = time.perf_counter()
    res = mm2._MapResult()
    offs = np.ascontiguousarray(roffs, dtype=np.uint64)
    rc = L.mm2_map_batch_device(ctx.h, gi.h, C.c_void_p(d_cat.data_ptr()), C.c_void_p(d_off.data_ptr()), offs.ctypes.data, offs.size - 1, C.byref(opts), C.byref(res))
    t1 = time.perf_counter()
    mr = mm2.MapResult(gi, res, offs.size - 1)
    t2 = time.perf_counter()
    tm = ctx.last_timings()
    t3 = time.perf_counter()
    mr.close()
    t4 = time.perf_counter()
    print("C call %.2f ms (host_call %.2f, records %.2f)  wrap %.2f  timings %.2f  close %.2f  total %.2f" % (1e3 * (t1 - t0), tm["host_call"], tm["host_records"], 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (t4 - t0)), flush=True)
