"""Scaled-down runs of BASELINE configs 3-5 (robustness at shape, GPU vs CPU oracle on a sample)."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minimap2_rs_b200 as mm2
from oracle import orc
from tools import gen

ap = argparse.ArgumentParser()
ap.add_argument("--which", default="c4,c5")
ap.add_argument("--scale", type=float, default=1.0)
a = ap.parse_args()
ctx = mm2.Context(0)
ncpu = os.cpu_count() or 1


def compare(gi, oi, cat, roffs, names, w, k, sample):
    t0 = time.perf_counter()
    res = ctx.map_batch(gi, cat, roffs, mm2.default_map_opts(w, k))
    t_gpu = time.perf_counter() - t0
    got = res.paf_lines(names)
    n = min(sample, len(names))
    t0 = time.perf_counter()
    want, st = oi.align_batch(cat[:int(roffs[n])], roffs[:n + 1], names[:n], orc.AlignOpts.default(w, k), threads=ncpu)
    t_cpu = time.perf_counter() - t0
    # first n reads' lines
    got_n = [l for l in got if int(l.split("\t")[0][1:]) < n]
    ok = got_n == want
    print("   reads=%d bases=%.2e  gpu %.3f s (%.2f Gbase/s)  cpu(%d thr) sample %d reads %.2f s  anchors/read %.0f cells/anchor %.1f rescued %d  PAF identical on sample: %s  timings %s"
          % (len(names), float(roffs[-1]), t_gpu, float(roffs[-1]) / t_gpu / 1e9, ncpu, n, t_cpu, st.n_anchors / max(1, n), st.cells / max(1, st.n_anchors),
             res.stats["n_rescued"], ok, {k_: round(v, 1) for k_, v in ctx.last_timings().items()}), flush=True)
    return ok


ok_all = True
if "c4" in a.which:
    print("C4-like: HiFi 15 kb reads, 0.5 % error, k=19 (map-hifi), multi-chromosome genome with N runs", flush=True)
    L = int(50e6 * a.scale)
    seqs, names = [], []
    for c in range(4):
        seqs.append(gen.genome(0xB2000003 + c, L, 1e-3 / 50, 50.0)); names.append("chr%d" % (c + 1))
        seqs.append(np.frombuffer(b"N", dtype=np.uint8)); names.append("pad%d" % c)
    offs = np.zeros(len(seqs) + 1, dtype=np.uint64); offs[1:] = np.cumsum([s.size for s in seqs])
    cat = np.concatenate(seqs)
    t0 = time.perf_counter(); gi = mm2.Index.build(ctx, cat, offs, names, w=10, k=19); print("   gpu index build %.3f s" % (time.perf_counter() - t0), gi.stats(), flush=True)
    oi = orc.Index.build(cat, offs, names, w=10, k=19, threads=ncpu)
    assert gi.stats() == oi.stats()
    nreads = int(20000 * a.scale)
    rc, ro = gen.reads(0xB2001004, cat, offs, nreads, 15000, 0.002, 0.0015, 0.0015)
    ok_all &= compare(gi, oi, rc, ro, ["r%d" % i for i in range(nreads)], 10, 19, 2000)
    gi.close()
if "c5" in a.which:
    print("C5-like: 100 kb ultra-long reads, 8 % error, repeat-rich genome (40 % tandem arrays, 20 % dispersed repeats)", flush=True)
    L = int(40e6 * a.scale)
    g = gen.repeat_genome(0xB2000005, L, 0.4, 0.2)
    offs = np.array([0, g.size], dtype=np.uint64)
    t0 = time.perf_counter(); gi = mm2.Index.build(ctx, g, offs, ["rep"]); print("   gpu index build %.3f s" % (time.perf_counter() - t0), gi.stats(), "mid_occ", gi.calc_mid_occ(), flush=True)
    oi = orc.Index.build(g, offs, ["rep"], threads=ncpu)
    assert gi.stats() == oi.stats() and gi.calc_mid_occ() == oi.calc_mid_occ()
    nreads = int(1000 * a.scale)
    rc, ro = gen.reads(0xB2001005, g, offs, nreads, 100000, 0.027, 0.027, 0.026)
    ok_all &= compare(gi, oi, rc, ro, ["r%d" % i for i in range(nreads)], 10, 15, 100)
    gi.close()
print("ALL OK" if ok_all else "MISMATCH")
