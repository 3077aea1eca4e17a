"""The C++ oracle (oracle/mm2_oracle.cpp) against tests/rust_transcript.py, an independent literal Python transcription
of the Rust sources.  Both were written from the reference text; agreement on the same seeded inputs is the pin on the
oracle that is available without a Rust toolchain.  Every stage is compared: minimizers (odd / even k,
HPC, N runs, short sequences), index (stats, calc_mid_occ, get), query filter, anchors (incl. odd rids), forward DP
(f, v, pprev and the inner-loop cell count), chains + scores (default and -n 1 general path, rescue), PAF text."""
import numpy as np
import pytest

import cases
import rust_transcript as rt


def _minis(lst):
    a = np.zeros(len(lst), dtype=[("key_span", "<u8"), ("rid_pos_strand", "<u8")])
    for i, (k, v) in enumerate(lst):
        a[i] = (k, v)
    return a


@pytest.mark.parametrize("w,k,hpc", [(10, 15, False), (10, 19, False), (10, 14, False), (10, 16, False), (3, 5, False), (1, 7, False),
                                     (5, 4, False), (10, 28, False), (19, 17, False), (10, 15, True), (10, 14, True), (4, 6, True)])
def test_sketch(orc, w, k, hpc):
    rng = np.random.default_rng(11)
    seqs = [s for _, s in cases.sketch_cases() if len(s) <= 3100]
    seqs += [b"AT" * 200 + cases.rnd_seq(rng, 200) + b"ACGT" * 60, b"ACGTTGCA" * 40, cases.rnd_seq(rng, 900, "ACGTN")]
    for i, s in enumerate(seqs):
        got = orc.sketch(s, w, k, rid=5, is_hpc=hpc)
        want = []
        rt.sketch_sequence(s, w, k, 5, hpc, want)
        want = _minis(want)
        assert got.size == want.size and (got == want).all(), (i, w, k, hpc)


def _small_genome(gen, seed, n=60_000):
    g = gen.repeat_genome(seed, n, 0.3, 0.2) if seed % 2 else gen.genome(seed, n, 1e-3, 20.0)
    c1, c2 = n // 2, n // 2 + n // 3
    return [bytes(g[:c1]), b"N", bytes(g[c1:c2]), b"NNNN", bytes(g[c2:])], ["chrA", "d1", "chrB", "d2", "chrC"]


@pytest.mark.parametrize("w,k,b,seed", [(10, 15, 14, 1), (10, 19, 10, 2), (5, 8, 6, 3)])
def test_index(orc, gen, w, k, b, seed):
    seqs, names = _small_genome(gen, seed)
    cat, offs = cases.cat_offs(seqs)
    oi = orc.Index.build(cat, offs, names, w=w, k=k, b=b, threads=2)
    ti = rt.Index.build(list(zip(names, seqs)), w, k, b, 0)
    assert oi.stats() == ti.stats()
    for frac in (2e-4, 1e-2, 0.5, 0.0, 1.0):
        assert oi.calc_mid_occ(frac) == ti.calc_mid_occ(frac), frac
    mv = orc.sketch(seqs[0][:8000], w, k)
    probes = [int(m["key_span"]) >> 8 for m in mv[::5]] + [0, 1, (1 << (2 * k)) - 1]
    for minier in probes:
        kind, occ = oi.get(minier)
        t = ti.get(minier)
        if t is None:
            assert kind == 0
        elif t[0] == "S":
            assert kind == 1 and int(occ[0]) == t[1]
        else:
            assert kind == 2 and occ.tolist() == list(t[1])


def _align_both(orc, oi, ti, cat, roffs, names, **kw):
    n = len(names)
    opts = orc.AlignOpts.default(kw.get("w", 10), kw.get("k", 15))
    for f in ("max_gap", "bw", "bw_long", "min_cnt", "min_chain_score", "mask_level", "pri_ratio", "best_n", "frac_top_repetitive"):
        if f in kw:
            setattr(opts, f, kw[f])
    got, st = oi.align_batch(cat, roffs, names, opts, threads=1)
    want = []
    cells = 0
    for r in range(n):
        q = bytes(cat[int(roffs[r]):int(roffs[r + 1])])
        stages = {}
        tkw = {a: kw[a] for a in ("w", "k", "frac_top_repetitive", "max_gap", "bw", "bw_long", "min_cnt", "min_chain_score", "mask_level",
                                  "pri_ratio", "best_n") if a in kw}
        want += rt.align_one(ti, names[r], q, stages=stages, **tkw)
        cells += stages["trace"].get("cells", 0) if stages["trace"] else 0
    return got, want, st, cells


def test_align_paf_default(orc, gen):
    seqs, names = _small_genome(gen, 4, 80_000)
    # even rids only carry sequence; reads drawn from them
    cat, offs = cases.cat_offs(seqs)
    oi = orc.Index.build(cat, offs, names, threads=2)
    ti = rt.Index.build(list(zip(names, seqs)), 10, 15, 14, 0)
    rc, roffs = gen.reads(21, cat, offs, 10, 1500, 0.03, 0.03, 0.03)
    rn = ["q%d" % i for i in range(10)]
    got, want, st, cells = _align_both(orc, oi, ti, rc, roffs, rn)
    assert got == want and len(got) >= 8
    if st.n_rescued == 0:
        assert st.cells == cells   # same inner-loop iteration count (lchain.rs:80) when no rescue rerun happened


def test_align_paf_repeats_and_chimeras(orc, gen):
    g = gen.repeat_genome(9, 70_000, 0.4, 0.2)
    offs = np.array([0, g.size], dtype=np.uint64)
    oi = orc.Index.build(g, offs, ["rep"], threads=2)
    ti = rt.Index.build([("rep", bytes(g))], 10, 15, 14, 0)
    rc, roffs = gen.reads(3, g, offs, 6, 2400, 0.02, 0.02, 0.02)
    # chimeric reads (two distant pieces) force rescue_long_join
    reads = [bytes(rc[int(roffs[i]):int(roffs[i + 1])]) for i in range(6)]
    reads.append(reads[0][:1200] + reads[3][:1200])
    reads.append(reads[1][:400] + cases.rnd_seq(np.random.default_rng(2), 1500) + reads[1][400:800])
    cat, ro = cases.cat_offs(reads)
    rn = ["c%d" % i for i in range(len(reads))]
    got, want, st, _ = _align_both(orc, oi, ti, cat, ro, rn)
    assert got == want
    assert st.n_rescued > 0


@pytest.mark.parametrize("kw", [dict(min_cnt=1, min_chain_score=10), dict(min_cnt=1, min_chain_score=1, best_n=2, pri_ratio=0.3, mask_level=0.9),
                                dict(bw=100, bw_long=3000, max_gap=800), dict(k=14, w=10), dict(k=19, w=10), dict(k=16, w=5, min_cnt=1, min_chain_score=5)])
def test_align_paf_options(orc, gen, kw):
    g = gen.repeat_genome(13, 60_000, 0.2, 0.2)
    offs = np.array([0, g.size], dtype=np.uint64)
    w, k = kw.get("w", 10), kw.get("k", 15)
    oi = orc.Index.build(g, offs, ["g"], w=w, k=k, threads=2)
    ti = rt.Index.build([("g", bytes(g))], w, k, 14, 0)
    rc, roffs = gen.reads(17, g, offs, 6, 1200, 0.02, 0.02, 0.02)
    rn = ["o%d" % i for i in range(6)]
    got, want, _, _ = _align_both(orc, oi, ti, rc, roffs, rn, **kw)
    assert got == want and got


def test_filter_and_anchors_stagewise(orc, gen):
    g = gen.repeat_genome(5, 50_000, 0.5, 0.1)
    seqs = [bytes(g[:30_000]), bytes(g[30_000:])]       # rid 1 is odd: sign-extended anchors (seeds.rs:64-71)
    cat, offs = cases.cat_offs(seqs)
    oi = orc.Index.build(cat, offs, ["e", "o"], threads=2)
    ti = rt.Index.build([("e", seqs[0]), ("o", seqs[1])], 10, 15, 14, 0)
    for lo in (1000, 31_000):
        q = bytes(g[lo:lo + 3000])
        mv_o = orc.sketch(q, 10, 15)
        mv_t = rt.collect_query_minimizers(q, 10, 15)
        assert (mv_o == _minis(mv_t)).all()
        kept_o = orc.filter_query_minimizers(mv_o)
        rt.filter_query_minimizers(mv_t, 10, 0.01)
        assert kept_o.size == len(mv_t) and (kept_o == _minis(mv_t)).all()
        for mid_occ in (10, 3, 2147483647):
            a_o = oi.anchors(kept_o, len(q), mid_occ)
            a_t = rt.build_anchors_filtered(ti, mv_t, len(q), mid_occ)
            assert a_o.size == len(a_t)
            assert a_o["x"].tolist() == [a[0] for a in a_t] and a_o["y"].tolist() == [a[1] for a in a_t]
            if lo == 1000 and mid_occ == 10 and a_o.size:
                p = orc.default_chain_params(15)
                o = orc.chain_dp_all(a_o, p)
                tr = {}
                chains, scores = rt.chain_dp_all(a_t, rt.default_chain_params(15), tr)
                assert o["f"].tolist() == tr["f"] and o["v"].tolist() == tr["v"] and o["pprev"].tolist() == tr["pprev"]
                assert o["cells"] == tr["cells"]
                assert [c.tolist() for c in o["chains"]] == chains and o["scores"].tolist() == scores
