"""A/B timing of library builds: `python tools/ab_step.py lib1.so lib2.so ...` runs the device-resident mapping step of
configs[1] (reduced: 50,000 reads) once per library in a fresh process and prints its stage times and a result checksum."""
import os
import subprocess
import sys

if len(sys.argv) > 1 and sys.argv[1] != "--child":
    for lib in sys.argv[1:]:
        env = dict(os.environ, MM2_LIB_PATH=os.path.abspath(lib))
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True)
        print(os.path.basename(lib), (r.stdout.strip().split("\n") or [""])[-1], r.stderr.strip()[-300:] if r.returncode else "", flush=True)
    sys.exit(0)

import hashlib

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minimap2_rs_b200 as mm2
from tools import gen

glen = 145_138_636
g = gen.genome(0xB2000002, glen)
goffs = np.array([0, glen], dtype=np.uint64)
cat, roffs = gen.reads(0xB2001002, g, goffs, 50_000, 10_000, 0.0333, 0.0333, 0.0333)
ctx = mm2.Context(0)
gi = mm2.Index.build(ctx, g, goffs, ["chr8"])
d_cat = torch.empty(cat.size + 64, dtype=torch.uint8, device="cuda")
d_cat[:cat.size].copy_(torch.from_numpy(cat))
d_off = torch.from_numpy(roffs.astype(np.int64)).cuda()
torch.cuda.synchronize()
acc, n = {}, 0
for it in range(8):
    res = ctx.map_batch(gi, None, roffs, mm2.default_map_opts(10, 15), device_ptrs=(d_cat.data_ptr(), d_off.data_ptr()))
    if it >= 3:
        for k, v in ctx.last_timings().items():
            acc[k] = acc.get(k, 0.0) + v
        n += 1
    if it == 7:
        sha = hashlib.sha256(np.ascontiguousarray(res.recs).tobytes()).hexdigest()[:12]
    res.close()
print({k: round(v / n, 3) for k, v in acc.items() if k in ("sketch", "lookup", "anchor_sort", "chain")}, sha)
