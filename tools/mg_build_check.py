"""torchrun script: bucket-partitioned multi-GPU index build over NCCL; rank 0 checks it against the single-GPU build.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/mg_build_check.py [--chroms 16] [--chrom-mbp 20] [--k 15]
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import minimap2_rs_b200 as mm2
from minimap2_rs_b200 import multi_gpu
from tools import gen

ap = argparse.ArgumentParser()
ap.add_argument("--chroms", type=int, default=16)
ap.add_argument("--chrom-mbp", type=float, default=20.0)
ap.add_argument("--k", type=int, default=15)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--no-check", action="store_true")
a = ap.parse_args()
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
# C3-shaped genome: chromosomes at even rids, 1-bp N records at odd rids, 0.1 % N runs
L = int(a.chrom_mbp * 1e6)
seqs, names = [], []
for c in range(a.chroms):
    seqs.append(gen.genome(0xB2000003 + c, L, 1e-3 / 50, 50.0))
    names.append("chr%d" % (c + 1))
    if c + 1 < a.chroms:
        seqs.append(np.frombuffer(b"N", dtype=np.uint8))
        names.append("pad%d" % (c + 1))
offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
offs[1:] = np.cumsum([s.size for s in seqs])
pin = mm2.PinnedBuffer(int(offs[-1]))
cat = pin.array(np.uint8, int(offs[-1]))
o = 0
for s in seqs:
    cat[o:o + s.size] = s
    o += s.size
del seqs
ctx = mm2.Context(lr)
times = []
for rep in range(a.reps):
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    gi = multi_gpu.build_index_sharded(ctx, cat, offs, names, w=10, k=a.k, b=14)
    torch.cuda.synchronize()
    dist.barrier()
    times.append(time.perf_counter() - t0)
    if rep + 1 < a.reps:
        gi.close()
t = torch.tensor([min(times)], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
out = {"n_gpus": world, "genome_bp": int(offs[-1]), "k": a.k, "sharded_build_s": float(t[0]), "gbp_per_s": int(offs[-1]) / float(t[0]) / 1e9,
       "stats": gi.stats()}
# every rank must hold the same replicated index
sig = torch.tensor([gi.stats()[0], gi.calc_mid_occ()], dtype=torch.int64, device="cuda")
sigs = [torch.empty_like(sig) for _ in range(world)]
dist.all_gather(sigs, sig)
out["replicas_agree"] = all(bool((s == sigs[0]).all()) for s in sigs)
if rank == 0 and not a.no_check:
    t0 = time.perf_counter()
    g1 = mm2.Index.build(ctx, cat, offs, names, w=10, k=a.k, b=14)
    out["single_gpu_build_s"] = time.perf_counter() - t0
    p1, p2 = "/tmp/mg_sharded.mmi", "/tmp/mg_single.mmi"
    gi.save_to_mmi(p1)
    g1.save_to_mmi(p2)
    h1 = hashlib.sha256(open(p1, "rb").read()).hexdigest()
    h2 = hashlib.sha256(open(p2, "rb").read()).hexdigest()
    out["mmi_identical_to_single_gpu"] = h1 == h2
    out["mmi_bytes"] = os.path.getsize(p1)
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
