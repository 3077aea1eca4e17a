"""BASELINE configs[3] at full size on ONE GPU: 3.1 Gbp synthetic genome (16 chromosomes at even rids, one-base N records at
odd rids, 0.1 % N runs), index k=19 w=10, 1 M simulated 15 kb HiFi-like reads (0.5 % error) mapped from pinned host memory."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minimap2_rs_b200 as mm2
from tools import gen

ap = argparse.ArgumentParser()
ap.add_argument("--chroms", type=int, default=16)
ap.add_argument("--chrom-bp", type=int, default=193_750_000)
ap.add_argument("--reads", type=int, default=1_000_000)
ap.add_argument("--read-len", type=int, default=15_000)
ap.add_argument("--out", default="gpurun_out/c4_full.json")
a = ap.parse_args()
t0 = time.perf_counter()
seqs, names = [], []
for c in range(a.chroms):
    seqs.append(gen.genome(0xB2000003 + c, a.chrom_bp, 1e-3 / 50, 50.0)); names.append("chr%d" % (c + 1))
    if c + 1 < a.chroms:
        seqs.append(np.frombuffer(b"N", dtype=np.uint8)); names.append("pad%d" % c)
offs = np.zeros(len(seqs) + 1, dtype=np.uint64); offs[1:] = np.cumsum([s.size for s in seqs])
cat = np.concatenate(seqs); del seqs
t_gen = time.perf_counter() - t0
ctx = mm2.Context(0)
t0 = time.perf_counter(); gi = mm2.Index.build(ctx, cat, offs, names, w=10, k=19); t_b1 = time.perf_counter() - t0
bt = gi.build_timings(); st = gi.stats()
t0 = time.perf_counter()
rc, ro = gen.reads(0xB2001004, cat, offs, a.reads, a.read_len, 0.002, 0.0015, 0.0015)
pin = mm2.PinnedBuffer(rc.size); pr = pin.array(np.uint8, rc.size); pr[:] = rc; del rc
t_reads = time.perf_counter() - t0
opts = mm2.default_map_opts(10, 19)
res = ctx.map_batch(gi, pr[: int(ro[20000])], ro[:20001], opts); res.close()       # warm-up: arenas, workers
t0 = time.perf_counter(); res = ctx.map_batch(gi, pr, ro, opts); t_map = time.perf_counter() - t0
stats = dict(res.stats); n_recs = int(res.n_recs)
tim = {k: round(v, 2) for k, v in ctx.last_timings().items()}
recs = res.recs.copy(); res.close()
# consistency: the first 3000 reads through the single-context path give the same records as inside the pipelined run
o2 = mm2.default_map_opts(10, 19); o2.want_stage_dump = 1
r2 = ctx.map_batch(gi, pr[: int(ro[3000])], ro[:3001], o2)
same = bool((r2.recs == recs[recs["read_id"] < 3000]).all()); r2.close()
cm = recs["cm"].astype(np.float64)
span = (recs["qend"].astype(np.int64) - recs["qstart"].astype(np.int64)) / recs["qlen"].astype(np.float64)
out = {"config": "configs[3]: %d x %d bp reads (0.5 %% error) vs %.2f Gbp / %d records, k=19 w=10, one B200" % (a.reads, a.read_len, cat.size / 1e9, len(names)),
       "genome_gen_s": t_gen, "reads_gen_and_pin_s": t_reads, "index_build_wall_s": t_b1, "index_build_device_ms": bt, "index_stats": st,
       "map_wall_s": t_map, "mapped_bases_per_s": float(ro[-1]) / t_map, "stage_ms_summed_over_subbatches": tim, "stats": stats, "paf_records": n_recs,
       "mean_chain_anchors": float(cm.mean()) if cm.size else 0.0, "median_query_span_fraction": float(np.median(span)) if span.size else 0.0,
       "first_3000_reads_identical_to_single_context_path": same}
print(json.dumps(out))
open(a.out, "w").write(json.dumps(out, indent=1))
