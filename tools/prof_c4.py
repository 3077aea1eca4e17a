"""Stage times of one device-resident mapping step on a configs[3]-shaped workload (HiFi 15 kb reads, k=19), single stream."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import minimap2_rs_b200 as mm2
from tools import gen
nreads = int(sys.argv[1]) if len(sys.argv) > 1 else 66667
seqs, names = [], []
for c in range(4):
    seqs.append(gen.genome(0xB2000003 + c, 50_000_000, 1e-3 / 50, 50.0)); names.append("chr%d" % (c + 1))
    seqs.append(np.frombuffer(b"N", dtype=np.uint8)); names.append("pad%d" % c)
offs = np.zeros(len(seqs) + 1, dtype=np.uint64); offs[1:] = np.cumsum([s.size for s in seqs])
cat = np.concatenate(seqs)
ctx = mm2.Context(0)
gi = mm2.Index.build(ctx, cat, offs, names, w=10, k=19)
rc, ro = gen.reads(0xB2001004, cat, offs, nreads, 15000, 0.002, 0.0015, 0.0015)
d_cat = torch.empty(rc.size + 64, dtype=torch.uint8, device="cuda"); d_cat[:rc.size].copy_(torch.from_numpy(rc))
d_off = torch.from_numpy(ro.astype(np.int64)).cuda()
opts = mm2.default_map_opts(10, 19)
for it in range(4):
    res = ctx.map_batch(gi, None, ro, opts, device_ptrs=(d_cat.data_ptr(), d_off.data_ptr()))
    t = ctx.last_timings()
    dev = sum(v for k, v in t.items() if not k.startswith("host_"))
    print("step %d: %.2f ms device (%.2f Gbase/s) %s %s" % (it, dev, float(ro[-1]) / dev / 1e6, {k: round(v, 2) for k, v in t.items()}, res.stats), flush=True)
    res.close()
