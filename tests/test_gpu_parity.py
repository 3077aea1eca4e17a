"""GPU parity suite (-m gpu): every stage of the CUDA path, called through the C ABI, against the CPU oracle — bit-exact."""
import os

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


def _eq(a, b, what):
    assert a.size == b.size, "%s: size %d vs %d" % (what, a.size, b.size)
    if a.size:
        bad = np.nonzero(a != b)[0]
        assert bad.size == 0, "%s: first mismatch at %d: %s vs %s" % (what, bad[0], a[bad[0]], b[bad[0]])


# ---- sketch ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("w,k", cases.WK)
def test_sketch_tile_kernel(ctx, orc, w, k):
    for name, s in cases.sketch_cases(big=(w, k) in ((10, 15), (10, 19))):
        _eq(ctx.sketch_sequence(s, w, k, rid=3), orc.sketch(s, w, k, rid=3), "sketch %s w=%d k=%d" % (name, w, k))


@pytest.mark.parametrize("w,k,hpc", [(10, 14, False), (10, 16, False), (5, 4, False), (10, 28, False), (10, 15, True), (10, 14, True)])
def test_sketch_literal_kernel(ctx, orc, w, k, hpc):
    rng = np.random.default_rng(4)
    seqs = [s for _, s in cases.sketch_cases()] + [b"AT" * 400 + cases.rnd_seq(rng, 300) + b"ACGT" * 100]
    for i, s in enumerate(seqs):
        _eq(ctx.sketch_sequence(s, w, k, rid=1, is_hpc=hpc), orc.sketch(s, w, k, rid=1, is_hpc=hpc), "literal %d" % i)


@pytest.mark.parametrize("w,k", [(10, 15), (10, 19), (11, 21), (19, 17), (9, 5), (64, 27)])
def test_sketch_hpc_tile_kernel(ctx, orc, w, k):
    """-H on the tile kernel (odd k, 9 <= w <= 64): same k-mers as the plain sketch, kmer_span = sum over the last k bases of
    the homopolymer run that remains from each base on (sketch.rs:51-61: the loop index is not advanced), k-mers with a span
    of 256 or more dropped; runs across tile and sequence boundaries, runs next to N, runs longer than 256"""
    rng = np.random.default_rng(14)
    seqs = [s for _, s in cases.sketch_cases()]
    seqs += [b"A" * 5000, cases.rnd_seq(rng, 2000) + b"C" * 300 + cases.rnd_seq(rng, 1800) + b"G" * 40 + b"N" + b"G" * 40 + cases.rnd_seq(rng, 500),
             cases.rnd_seq(rng, 6000, "AAAC"), cases.rnd_seq(rng, 4096, "AACCGGTTN"), b"T" * 255 + b"A" + b"T" * 256 + cases.rnd_seq(rng, 100)]
    for i, s in enumerate(seqs):
        _eq(ctx.sketch_sequence(s, w, k, rid=2, is_hpc=True), orc.sketch(s, w, k, rid=2, is_hpc=True), "hpc tile %d w=%d k=%d" % (i, w, k))
    cat, offs = cases.cat_offs(seqs)
    mv, mo = ctx.sketch_batch(cat, offs, w, k, rid_base=0, rid_step=1, is_hpc=True)
    for i, s in enumerate(seqs):
        _eq(mv[int(mo[i]):int(mo[i + 1])], orc.sketch(s, w, k, rid=i, is_hpc=True), "hpc batch seq %d" % i)


def test_sketch_batch_many_sequences(ctx, orc):
    rng = np.random.default_rng(8)
    seqs = [cases.rnd_seq(rng, int(n)) for n in rng.integers(1, 9000, 300)] + [b"N" * 100, b"A"]
    cat, offs = cases.cat_offs(seqs)
    mv, mo = ctx.sketch_batch(cat, offs, 10, 15, rid_base=0, rid_step=1)
    for i, s in enumerate(seqs):
        _eq(mv[int(mo[i]):int(mo[i + 1])], orc.sketch(s, 10, 15, rid=i), "seq %d" % i)
    mv0, _ = ctx.sketch_batch(cat, offs, 10, 19)
    want = np.concatenate([orc.sketch(s, 10, 19, rid=0) for s in seqs])
    _eq(mv0, want, "rid 0 batch")


def test_sketch_concurrent_contexts(mm2, orc, gen):
    """four contexts sketch the same batch at the same time (what the pipelined mapping path does): the kernels of different
    launches share the SMs, so that a CTA runs with fewer resident neighbours and other timings than alone"""
    import threading
    g = gen.genome(52, 1_000_000)
    offs = np.array([0, g.size], dtype=np.uint64)
    cat, roffs = gen.reads(9, g, offs, 1500, 4000, 0.03, 0.03, 0.03)
    ctxs = [mm2.Context(0) for _ in range(4)]
    want, wo = ctxs[0].sketch_batch(cat, roffs, 10, 15)
    for i in (0, 7, 1499):
        _eq(want[int(wo[i]):int(wo[i + 1])], orc.sketch(bytes(cat[int(roffs[i]):int(roffs[i + 1])]), 10, 15, rid=0), "read %d" % i)
    bad = []

    def work(c):
        for rep in range(6):
            mv, mo = c.sketch_batch(cat, roffs, 10, 15)
            if mv.size != want.size or not (mv == want).all() or not (mo == wo).all():
                bad.append(rep)

    th = [threading.Thread(target=work, args=(c,)) for c in ctxs]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not bad


def test_sketch_rejects_what_the_reference_asserts_on(ctx, mm2):
    for args in ((b"", 10, 15), (b"ACGT", 0, 15), (b"ACGT", 256, 15), (b"ACGT", 10, 0), (b"ACGT", 10, 29)):
        with pytest.raises(mm2.Mm2Error) as e:
            ctx.sketch_sequence(*args)
        assert e.value.code == mm2.MM2_E_ARG


# ---- index ----------------------------------------------------------------------------------------------------
def _genome_records(gen, seed, total, with_n=True):
    g = gen.repeat_genome(seed, total, 0.3, 0.2) if seed % 2 else gen.genome(seed, total, 1e-3 if with_n else 0.0, 30.0)
    cut1, cut2 = total // 2, total // 2 + total // 3
    seqs = [bytes(g[:cut1]), b"N", bytes(g[cut1:cut2]), b"NNNNNNNN", bytes(g[cut2:])]
    return seqs, ["chrA", "dummy1", "chrB extra words", "d2", "chrC"]


@pytest.mark.parametrize("w,k,b,seed", [(10, 15, 14, 1), (10, 19, 14, 2), (10, 15, 8, 3), (5, 5, 14, 4), (11, 21, 12, 5), (10, 14, 10, 6)])
def test_index_build_mmi_bytes(ctx, mm2, orc, gen, tmp_path, w, k, b, seed):
    seqs, names = _genome_records(gen, seed, 400_000)
    cat, offs = cases.cat_offs(seqs)
    gi = mm2.Index.build(ctx, cat, offs, [n.split()[0] for n in names], w=w, k=k, b=b)
    oi = orc.Index.build(cat, offs, [n.split()[0] for n in names], w=w, k=k, b=b, threads=4)
    assert gi.stats() == oi.stats()
    for frac in (2e-4, 1e-2, 0.5, 0.0, 1.0):
        assert gi.calc_mid_occ(frac) == oi.calc_mid_occ(frac)
    pg, po = str(tmp_path / "g.mmi"), str(tmp_path / "o.mmi")
    gi.save_to_mmi(pg)
    oi.save_mmi(po)
    bg, bo = open(pg, "rb").read(), open(po, "rb").read()
    assert len(bg) == len(bo)
    assert bg == bo, "first differing byte %d" % next(i for i in range(len(bg)) if bg[i] != bo[i])
    # lookups
    mv = orc.sketch(seqs[0][:20000], w, k)
    for m in mv[::37]:
        minier = int(m["key_span"]) >> 8
        kg, og = gi.get(minier)
        ko, oo = oi.get(minier)
        assert kg == ko and (og == oo).all()
    assert gi.get((1 << (2 * k)) - 1 if k < 28 else 12345)[0] == oi.get((1 << (2 * k)) - 1 if k < 28 else 12345)[0]
    assert gi.get_ref_subseq(0, 100, 180) == bytes(seqs[0][100:180]).upper().replace(b"R", b"N")
    assert gi.seq(2) == ("chrB", len(seqs[2]))


def test_index_load_roundtrips(ctx, mm2, orc, gen, tmp_path):
    seqs, names = _genome_records(gen, 7, 300_000)
    cat, offs = cases.cat_offs(seqs)
    oi = orc.Index.build(cat, offs, names, threads=4)
    po, pn = str(tmp_path / "o.mmi"), str(tmp_path / "o.idx")
    oi.save_mmi(po)
    oi.save_native(pn)
    g1 = mm2.Index.load_from_mmi(ctx, po)
    g2 = mm2.Index.load_from_file(ctx, pn)
    g3 = mm2.Index.load_auto(ctx, po)
    for g in (g1, g2, g3):
        assert g.stats() == oi.stats() and g.calc_mid_occ() == oi.calc_mid_occ()
        assert (g.w, g.k, g.b, g.n_seq) == (10, 15, 14, 5)
    p1, p2 = str(tmp_path / "g1.mmi"), str(tmp_path / "g2.idx")
    g1.save_to_mmi(p1)
    g2.save_to_file(p2)
    assert open(p1, "rb").read() == open(po, "rb").read()
    assert open(p2, "rb").read() == open(pn, "rb").read()
    with pytest.raises(mm2.Mm2Error) as e:
        mm2.Index.load_from_mmi(ctx, pn)
    assert e.value.code == mm2.MM2_E_FORMAT
    # FASTA ingest + load_auto fallback (main.rs:135-145)
    fa = str(tmp_path / "ref.fa")
    with open(fa, "wb") as f:
        for n, s in zip(names, seqs):
            f.write(b">" + n.encode() + b"\n")
            for i in range(0, len(s), 70):
                f.write(s[i:i + 70] + b"\n")
    g4 = mm2.Index.load_auto(ctx, fa)
    o4 = orc.Index.build_fasta(fa)
    p4, q4 = str(tmp_path / "g4.mmi"), str(tmp_path / "o4.mmi")
    g4.save_to_mmi(p4)
    o4.save_mmi(q4)
    assert open(p4, "rb").read() == open(q4, "rb").read()


# ---- seeds / chain stage entry points ------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def small_world(ctx, mm2, orc, gen):
    g = gen.repeat_genome(77, 300_000, 0.4, 0.2)
    offs = np.array([0, g.size], dtype=np.uint64)
    gi = mm2.Index.build(ctx, g, offs, ["c"])
    oi = orc.Index.build(g, offs, ["c"], threads=4)
    cat, roffs = gen.reads(5, g, offs, 12, 3000, 0.02, 0.02, 0.02)
    return g, gi, oi, cat, roffs


def test_filter_and_anchors(ctx, orc, small_world):
    g, gi, oi, cat, roffs = small_world
    mid = max(10, oi.calc_mid_occ())
    assert max(10, gi.calc_mid_occ()) == mid
    for r in range(12):
        q = cat[int(roffs[r]):int(roffs[r + 1])]
        mv = orc.sketch(q, 10, 15)
        for qmax, frac in ((10, 0.01), (2, 0.001), (10, 0.0)):
            _eq(ctx.filter_query_minimizers(mv, qmax, frac), orc.filter_query_minimizers(mv, qmax, frac), "filter r%d" % r)
        mvf = orc.filter_query_minimizers(mv)
        for mo in (mid, 200, 0x7fffffff):
            _eq(ctx.build_anchors_filtered(gi, mvf, q.size, mo), oi.anchors(mvf, q.size, mo), "anchors r%d mid_occ %d" % (r, mo))
    # a read made of one repeated unit: the query-side filter must drop its over-represented minimizers
    unit = bytes(g[1000:1180])
    q = np.frombuffer(unit * 30, dtype=np.uint8)
    mv = orc.sketch(q, 10, 15)
    a, b = ctx.filter_query_minimizers(mv), orc.filter_query_minimizers(mv)
    assert b.size < mv.size
    _eq(a, b, "repeat read filter")


def test_chain_dp_all(ctx, orc, small_world):
    g, gi, oi, cat, roffs = small_world
    for r in range(12):
        q = cat[int(roffs[r]):int(roffs[r + 1])]
        a = oi.anchors(orc.filter_query_minimizers(orc.sketch(q, 10, 15)), q.size, 200 if r % 2 else 60)
        if a.size > 30000:
            a = a[:30000]
        for variant in range(4):
            p = orc.default_chain_params(15)
            if variant == 1:
                p.bw = 20000
            if variant == 2:
                p.max_chain_skip, p.max_chain_iter = 3, 50
            if variant == 3:
                p.min_cnt, p.min_chain_score = 1, 10
            o = orc.chain_dp_all(a, p)
            import ctypes as C
            import minimap2_rs_b200 as m
            gp = m.ChainParams.from_buffer_copy(bytes(p))
            gres = ctx.chain_dp_all(a, gp)
            _eq(gres["f"], o["f"], "f r%d v%d" % (r, variant))
            _eq(gres["pprev"], o["pprev"], "pprev r%d v%d" % (r, variant))
            _eq(gres["v"], o["v"], "v r%d v%d" % (r, variant))
            assert len(gres["chains"]) == len(o["chains"])
            _eq(gres["scores"], o["scores"], "scores")
            for cg, co in zip(gres["chains"], o["chains"]):
                _eq(cg, co, "chain")


@pytest.mark.parametrize("density,max_skip,max_iter", [(5, 25, 5000), (40, 25, 5000), (40, 0, 5000), (40, 2, 31), (40, 25, 32), (40, 1, 33),
                                                      (200, 25, 5000), (200, 3, 64), (200, 25, 1), (1000, 25, 5000), (1000, 5, 700)])
@pytest.mark.parametrize("which", ["warp_per_read", "cta_per_read"])
def test_chain_dp_all_synthetic_windows(ctx, dense_ctx, orc, density, max_skip, max_iter, which):
    """anchors placed directly (no index): predecessor windows around and beyond the 32-slot register ring of the kernel,
    collinear runs (marks -> max_chain_skip breaks), ties, both strands, several rids"""
    import minimap2_rs_b200 as m
    if which == "cta_per_read":
        ctx = dense_ctx        # n >= 4096 would take that kernel anyway; the other arm keeps n below it
    rng = np.random.default_rng(density * 1000 + max_skip * 10 + max_iter)
    n = 6000 if which == "cta_per_read" else 3600
    span = 15
    parts = []
    for rid, rev, cnt in ((0, 0, n // 2), (0, 1, n // 6), (2, 0, n // 6), (4, 1, n // 6)):
        extent = cnt * 5000 // density          # ~density anchors per max_dist_x of target
        rpos = np.sort(rng.integers(20, 20 + extent, cnt)).astype(np.int64)
        diag = rng.choice([0, 0, 0, 7, -13, 400, -2500], cnt).astype(np.int64)
        run = rng.integers(0, 4, cnt) == 0      # exact collinear repeats of the previous anchor's diagonal
        diag[1:][run[1:]] = diag[:-1][run[1:]]
        qpos = np.clip(rpos + diag + rng.integers(-3, 4, cnt) * (rng.integers(0, 3, cnt) == 0), 14, None)
        x = (np.uint64(rev) << np.uint64(63)) | (np.uint64(rid) << np.uint64(32)) | rpos.astype(np.uint64)
        y = (np.uint64(span) << np.uint64(32)) | qpos.astype(np.uint64)
        parts.append(np.stack([x, y], axis=1))
    xy = np.concatenate(parts)
    order = np.lexsort((xy[:, 1], xy[:, 0]))
    xy = xy[order]
    a = np.zeros(xy.shape[0], dtype=orc.ANCHOR_DT)
    a["x"], a["y"] = xy[:, 0], xy[:, 1]
    p = orc.default_chain_params(15)
    p.max_chain_skip, p.max_chain_iter = max_skip, max_iter
    for bw in (p.bw, 100000):
        p.bw = bw
        o = orc.chain_dp_all(a, p)
        gres = ctx.chain_dp_all(a, m.ChainParams.from_buffer_copy(bytes(p)))
        _eq(gres["f"], o["f"], "f")
        _eq(gres["pprev"], o["pprev"], "pprev")
        _eq(gres["v"], o["v"], "v")
        _eq(gres["scores"], o["scores"], "scores")
        assert len(gres["chains"]) == len(o["chains"])
        for cg, co in zip(gres["chains"], o["chains"]):
            _eq(cg, co, "chain")


# ---- the batched mapping path: PAF lines ---------------------------------------------------------------------------
def _map_compare(ctx, mm2, orc, gi, oi, cat, roffs, names, opts_w_k=(10, 15), dump=True):
    w, k = opts_w_k
    o = mm2.default_map_opts(w, k)
    o.want_stage_dump = 1 if dump else 0
    res = ctx.map_batch(gi, cat, roffs, o)
    oo = orc.AlignOpts.default(w, k)
    want, st = oi.align_batch(cat, roffs, names, oo, threads=8)
    if dump:
        sd = res.stage
        mid = max(10, oi.calc_mid_occ())
        for r in range(len(names)):
            q = cat[int(roffs[r]):int(roffs[r + 1])]
            mv = orc.sketch(q, w, k)
            gm = sd["minis"][int(sd["mini_offs"][r]):int(sd["mini_offs"][r + 1])]
            _eq(gm, mv, "minimizers of read %d" % r)
            keep = sd["mini_keep"][int(sd["mini_offs"][r]):int(sd["mini_offs"][r + 1])].astype(bool)
            mvf = orc.filter_query_minimizers(mv)
            _eq(gm[keep], mvf, "kept minimizers of read %d" % r)
            a = oi.anchors(mvf, q.size, mid)
            ga = sd["anchors"][int(sd["anchor_offs"][r]):int(sd["anchor_offs"][r + 1])]
            _eq(ga, a, "anchors of read %d" % r)
        assert res.stats["n_minimizers"] == st.n_minimizers and res.stats["n_anchors"] == st.n_anchors
        assert res.stats["n_minimizers_kept"] == st.n_minimizers_kept
    got = res.paf_lines(names)
    assert res.stats["n_rescued"] == st.n_rescued
    assert len(got) == len(want)
    for i, (x, y) in enumerate(zip(got, want)):
        assert x == y, "PAF line %d:\n%s\n%s" % (i, x, y)
    return res, st


def test_map_batch_random_genome(ctx, mm2, orc, gen):
    g = gen.genome(0xB2000001, 3_000_000)
    offs = np.array([0, g.size], dtype=np.uint64)
    gi = mm2.Index.build(ctx, g, offs, ["chr8"])
    oi = orc.Index.build(g, offs, ["chr8"], threads=8)
    for seed, n, ln, e in ((1, 200, 10000, 0.0333), (2, 300, 600, 0.002), (3, 50, 15000, 0.0017), (4, 40, 30000, 0.05)):
        cat, roffs = gen.reads(seed, g, offs, n, ln, e, e, e)
        names = ["r%06d" % i for i in range(n)]
        _map_compare(ctx, mm2, orc, gi, oi, cat, roffs, names)


def test_map_batch_ragged_and_empty_reads(ctx, mm2, orc, gen):
    g = gen.genome(5, 1_000_000, 1e-3, 40.0)
    offs = np.array([0, g.size], dtype=np.uint64)
    gi = mm2.Index.build(ctx, g, offs, ["ref"])
    oi = orc.Index.build(g, offs, ["ref"], threads=8)
    rng = np.random.default_rng(12)
    reads = []
    for i in range(120):
        ln = int(rng.choice([1, 10, 24, 25, 30, 100, 500, 2037, 2038, 2039, 5000, 12000]))
        p = int(rng.integers(0, g.size - ln))
        s = bytearray(g[p:p + ln].tobytes())
        if i % 7 == 0:
            s = bytearray(cases.rnd_seq(rng, ln))          # unrelated read: usually no anchors -> no PAF line
        if i % 5 == 0 and ln > 50:
            s = bytearray(bytes(s).translate(bytes.maketrans(b"ACGT", b"TGCA"))[::-1])
        reads.append(bytes(s))
    cat, roffs = cases.cat_offs(reads)
    _map_compare(ctx, mm2, orc, gi, oi, cat, roffs, ["q%d" % i for i in range(len(reads))])
    # an empty batch is fine
    res = ctx.map_batch(gi, np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.uint64))
    assert res.recs.size == 0


def test_map_batch_repeats_rescue_and_hifi_preset(ctx, mm2, orc, gen):
    g = gen.repeat_genome(77, 2_000_000, 0.4, 0.2)
    offs = np.array([0, g.size], dtype=np.uint64)
    w, k = mm2.apply_preset("map-hifi", 10, 15)
    gi = mm2.Index.build(ctx, g, offs, ["rep"], w=w, k=k)
    oi = orc.Index.build(g, offs, ["rep"], w=w, k=k, threads=8)
    cat, roffs = gen.reads(21, g, offs, 60, 8000, 0.002, 0.0015, 0.0015)
    # chimeric reads (two distant pieces) force rescue_long_join's rerun
    chim = []
    rng = np.random.default_rng(1)
    for i in range(20):
        a, b = int(rng.integers(0, g.size - 5000)), int(rng.integers(0, g.size - 5000))
        chim.append(g[a:a + 4000].tobytes() + g[b:b + 3000].tobytes())
    cat2, roffs2 = cases.cat_offs([cat[int(roffs[i]):int(roffs[i + 1])].tobytes() for i in range(60)] + chim)
    res, st = _map_compare(ctx, mm2, orc, gi, oi, cat2, roffs2, ["h%d" % i for i in range(80)], (w, k))
    assert st.n_rescued > 0


def test_map_batch_anchor_classes_and_exact_filter(ctx, mm2, orc, gen):
    """one batch that exercises every anchor-count class of the seeding stage (fused fill + merge sort up to 1024 and up to 4096
    anchors, device-built lists for the shared-memory and global-memory bitonic classes above) and the exact query filter
    behind the count sketch (reads listed by seed_hits_kernel<0>, filtered by filter_list_kernel, looked up by <1>)"""
    g = gen.repeat_genome(79, 2_000_000, 0.4, 0.2)
    offs = np.array([0, g.size], dtype=np.uint64)
    gi = mm2.Index.build(ctx, g, offs, ["rep"])
    oi = orc.Index.build(g, offs, ["rep"], threads=8)
    reads = []
    for seed, n, ln in ((31, 30, 700), (32, 30, 3000), (33, 20, 9000), (34, 8, 30000)):
        c, ro = gen.reads(seed, g, offs, n, ln, 0.02, 0.02, 0.02)
        reads += [c[int(ro[i]):int(ro[i + 1])].tobytes() for i in range(n)]
    reads += [b"AC" * 4000, b"ACGTTGCA" * 1500]                     # low-complexity reads: the same few k-mers hundreds of times
    cat, roffs = cases.cat_offs(reads)
    res, st = _map_compare(ctx, mm2, orc, gi, oi, cat, roffs, ["c%d" % i for i in range(len(reads))])
    na = np.diff(res.stage["anchor_offs"].astype(np.int64))
    assert (na <= 1024).any() and ((na > 1024) & (na <= 4096)).any() and ((na > 4096) & (na <= 12288)).any() and (na > 12288).any(), np.sort(na)
    assert st.n_minimizers_kept < st.n_minimizers                      # the exact filter dropped something


def test_map_batch_cta_per_read_chaining(dense_ctx, mm2, orc, gen):
    """the whole mapping path with every read chained by chain_dense_kernel (repeat-rich genome: long windows, rescue)"""
    g = gen.repeat_genome(78, 1_500_000, 0.4, 0.2)
    offs = np.array([0, g.size], dtype=np.uint64)
    gi = mm2.Index.build(dense_ctx, g, offs, ["rep"])
    oi = orc.Index.build(g, offs, ["rep"], threads=8)
    cat, roffs = gen.reads(22, g, offs, 40, 12000, 0.02, 0.02, 0.02)
    rng = np.random.default_rng(2)
    chim = []
    for i in range(6):   # chimeric reads force rescue_long_join's rerun (second DP pass of the same CTA)
        a, b = int(rng.integers(0, g.size - 5000)), int(rng.integers(0, g.size - 5000))
        chim.append(g[a:a + 4000].tobytes() + g[b:b + 1500].tobytes())
    reads = [cat[int(roffs[i]):int(roffs[i + 1])].tobytes() for i in range(40)] + chim + [b"ACGT" * 10]
    cat2, roffs2 = cases.cat_offs(reads)
    res, st = _map_compare(dense_ctx, mm2, orc, gi, oi, cat2, roffs2, ["d%d" % i for i in range(len(reads))], dump=False)
    assert st.n_rescued > 0


def test_map_batch_multi_sequence_even_rids_and_panic(ctx, mm2, orc, gen):
    g = gen.genome(9, 1_200_000)
    seqs = [g[:500_000].tobytes(), b"N", g[500_000:900_000].tobytes(), g[900_000:].tobytes()]
    cat, offs = cases.cat_offs(seqs)
    names = ["c0", "n1", "c2", "c3odd"]
    gi = mm2.Index.build(ctx, cat, offs, names)
    oi = orc.Index.build(cat, offs, names, threads=8)
    reads = [seqs[0][1000:4000], seqs[2][5000:9000], seqs[3][100:3100], seqs[2][100_000:101_000]]
    rc, ro = cases.cat_offs(reads)
    res, st = _map_compare(ctx, mm2, orc, gi, oi, rc, ro, ["a", "b", "c", "d"], dump=False)
    assert st.n_panic == 1 and res.panic_reads.tolist() == [2]     # odd rid -> the reference panics (F5)


def test_map_batch_device_resident_and_stream(ctx, mm2, orc, gen):
    import torch
    g = gen.genome(31, 1_000_000)
    offs = np.array([0, g.size], dtype=np.uint64)
    gi = mm2.Index.build(ctx, g, offs, ["t"])
    oi = orc.Index.build(g, offs, ["t"], threads=8)
    cat, roffs = gen.reads(3, g, offs, 100, 5000, 0.02, 0.02, 0.02)
    d_cat = torch.from_numpy(cat.copy()).cuda()
    d_off = torch.from_numpy(roffs.astype(np.int64)).cuda()
    c2 = mm2.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    res = c2.map_batch(gi, None, roffs, device_ptrs=(d_cat.data_ptr(), d_off.data_ptr()))
    want, _ = oi.align_batch(cat, roffs, ["x%d" % i for i in range(100)], threads=8)
    assert res.paf_lines(["x%d" % i for i in range(100)]) == want
    assert c2.launch_count > 0 and "chain" in c2.last_timings()
    c2.close()


def test_gpu_matches_golden_fixtures(ctx, mm2, gen, tmp_path):
    """the committed vectors (tests/golden), without recomputing the oracle"""
    import hashlib
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    d = np.load(os.path.join(gold, "sketch.npz"))
    for i in range(int(d["n"])):
        mv = ctx.sketch_sequence(d["seq%d" % i].tobytes(), int(d["w"][i]), int(d["k"][i]), rid=1)
        assert (mv["key_span"] == d["key%d" % i]).all() and (mv["rid_pos_strand"] == d["val%d" % i]).all(), i
    g = gen.genome(0xB2000001, 200_000)
    offs = np.array([0, g.size], dtype=np.uint64)
    gi = mm2.Index.build(ctx, g, offs, ["chr8"])
    p = str(tmp_path / "g.mmi")
    gi.save_to_mmi(p)
    sha, size = open(os.path.join(gold, "index_200k.sha256")).read().split()
    assert hashlib.sha256(open(p, "rb").read()).hexdigest() == sha and os.path.getsize(p) == int(size)
    cat, roffs = gen.reads(0xB2001001, g, offs, 24, 3000, 0.02, 0.02, 0.02)
    res = ctx.map_batch(gi, cat, roffs)
    assert res.paf_lines(["r%02d" % i for i in range(24)]) == open(os.path.join(gold, "align_24reads.paf")).read().split("\n")[:-1]
    s = np.load(os.path.join(gold, "read0_stages.npz"))
    a = np.zeros(s["ax"].size, dtype=mm2.ANCHOR_DT)
    a["x"], a["y"] = s["ax"], s["ay"]
    r = ctx.chain_dp_all(a, mm2.default_chain_params(15))
    assert (r["f"] == s["f"]).all() and (r["pprev"] == s["pprev"]).all() and (r["chains"][0] == s["chain0"]).all()
    assert r["scores"][0] == int(s["score0"])


def test_cli_index_align_anchors_chain(ctx, mm2, orc, gen, tmp_path):
    """the `mm2rs` binary: same subcommands / flags / stdout as src/main.rs (Appendix D of SURVEY.md)"""
    import subprocess
    exe = os.path.join(os.path.dirname(mm2.LIB_PATH), "mm2rs")
    g = gen.genome(41, 600_000)
    seqs = [g[:350_000].tobytes(), b"N", g[350_000:].tobytes()]
    fa, qa = str(tmp_path / "ref.fa"), str(tmp_path / "q.fa")
    with open(fa, "wb") as f:
        for n, s in zip(["chrA desc", "gap", "chrB"], seqs):
            f.write(b">" + n.encode() + b"\n")
            for i in range(0, len(s), 60):
                f.write(s[i:i + 60] + b"\n")
    cat, offs = cases.cat_offs(seqs)
    goffs = np.array([0, g.size], dtype=np.uint64)
    rc, ro = gen.reads(4, g[:350_000], np.array([0, 350_000], dtype=np.uint64), 3, 4000, 0.03, 0.03, 0.03)
    with open(qa, "wb") as f:
        for i in range(3):
            f.write(b">read%d some comment\n" % i + rc[int(ro[i]):int(ro[i + 1])].tobytes() + b"\n")
    oi = orc.Index.build(cat, offs, ["chrA", "gap", "chrB"], threads=4)
    mmi = str(tmp_path / "ref.mmi")
    out = subprocess.run([exe, "index", "-d", mmi, fa], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    nk, ao, sp, tl = oi.stats()
    assert out.stdout == "kmer size: 15; skip: 10; is_hpc: 0; #seq: 3\ndistinct minimizers: %d (avg occ %.2f) avg spacing %.3f total length %d\n" % (nk, ao, sp, tl)
    po = str(tmp_path / "o.mmi")
    oi.save_mmi(po)
    assert open(mmi, "rb").read() == open(po, "rb").read()
    names = ["read%d" % i for i in range(3)]
    want, _ = oi.align_batch(rc, ro, names)
    # strict drop-in: first record only (main.rs:92-103)
    out = subprocess.run([exe, "align", mmi, qa], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout == want[0] + "\n", (out.stdout, out.stderr)
    # all reads, from the FASTA reference directly (load_index_auto fallback), into a file
    pf = str(tmp_path / "o.paf")
    out = subprocess.run([exe, "align", fa, qa, "--all-reads", "-o", pf, "-x", "map-ont"], capture_output=True, text=True)
    assert out.returncode == 0 and open(pf).read() == "\n".join(want) + "\n", out.stderr
    # extension: the same reads as FASTQ (multi-line sequence, '@' and '+' inside the qualities) give the same PAF
    qq = str(tmp_path / "q.fq")
    with open(qq, "wb") as f:
        for i in range(3):
            sq = rc[int(ro[i]):int(ro[i + 1])].tobytes()
            f.write(b"@read%d some comment\n" % i + sq[:1000] + b"\n" + sq[1000:] + b"\n+\n" + b"@+" * (len(sq) // 2) + b"I" * (len(sq) % 2) + b"\n")
    out = subprocess.run([exe, "align", mmi, qq, "--all-reads"], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout == "\n".join(want) + "\n", out.stderr
    out = subprocess.run([exe, "align", mmi, qq], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout == want[0] + "\n"
    # anchors / chain debug subcommands
    q0 = rc[int(ro[0]):int(ro[1])]
    a = oi.anchors(orc.filter_query_minimizers(orc.sketch(q0, 10, 15)), q0.size, max(10, oi.calc_mid_occ()))
    exp = "anchors: %d\n" % a.size + "".join("x=0x%016x y=0x%016x\n" % (int(v["x"]), int(v["y"])) for v in a[:10])
    out = subprocess.run([exe, "anchors", mmi, qa], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout == exp
    p = orc.default_chain_params(15)
    p.bw = 300
    ch = orc.chain_dp_all(a, p)["chains"][0]
    exp = "best_chain_len: %d\nstart: x=0x%016x y=0x%016x\nend:   x=0x%016x y=0x%016x\n" % (
        ch.size, int(a[ch[0]]["x"]), int(a[ch[0]]["y"]), int(a[ch[-1]]["x"]), int(a[ch[-1]]["y"]))
    out = subprocess.run([exe, "chain", mmi, qa, "-r", "300"], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout == exp, (out.stdout, exp)
    # errors: missing file -> "Error: ..." exit 1
    out = subprocess.run([exe, "align", str(tmp_path / "missing.mmi"), qa], capture_output=True, text=True)
    assert out.returncode == 1 and out.stderr.startswith("Error:")


def test_map_batch_packed_equals_ascii(mm2, orc, gen):
    """mm2_map_batch_packed (2-bit reads + positions of the non-ACGT bases) == mm2_map_batch on the ASCII reads, on the single
    path and on the pipelined path (sub-batch boundaries inside 16-base words, N and IUPAC letters next to them)"""
    g = gen.genome(53, 1_500_000)
    offs = np.array([0, g.size], dtype=np.uint64)
    cat, roffs = gen.reads(10, g, offs, 2500, 3001, 0.03, 0.03, 0.03)      # odd read length: reads start inside 16-base words
    cat = cat.copy()
    rng = np.random.default_rng(6)
    for p in rng.integers(0, cat.size, 4000):
        cat[p] = rng.choice(np.frombuffer(b"NnRYacgt", dtype=np.uint8))
    for lo in rng.integers(0, cat.size - 200, 30):
        cat[lo:lo + int(rng.integers(1, 120))] = ord("N")
    packed, n_pos = mm2.pack_reads(cat)
    assert n_pos.size > 3000 and (np.diff(n_pos.astype(np.int64)) > 0).all()
    for sub in (None, "1"):
        if sub:
            os.environ["MM2_SUBBATCH_MB"] = sub
        c = mm2.Context(0)
        os.environ.pop("MM2_SUBBATCH_MB", None)
        gi = mm2.Index.build(c, g, offs, ["p"])
        r1 = c.map_batch(gi, cat, roffs)
        r2 = c.map_batch_packed(gi, packed, n_pos, roffs)
        assert r1.recs.size == r2.recs.size and r1.recs.size > 2000 and (r1.recs == r2.recs).all(), sub
        assert r1.stats == r2.stats
    with pytest.raises(mm2.Mm2Error) as e:
        c.map_batch_packed(gi, packed, n_pos, roffs[1:])                    # offs[0] = 3001: not on a 16-base boundary
    assert e.value.code == mm2.MM2_E_ARG


def test_map_batch_pipelined_equals_single(mm2, orc, gen):
    """large host batches are split into sub-batches over two worker contexts; records must not change"""
    g = gen.genome(51, 2_000_000)
    offs = np.array([0, g.size], dtype=np.uint64)
    os.environ["MM2_SUBBATCH_MB"] = "2"
    c = mm2.Context(0)
    del os.environ["MM2_SUBBATCH_MB"]
    gi = mm2.Index.build(c, g, offs, ["p"])
    cat, roffs = gen.reads(8, g, offs, 3000, 4000, 0.03, 0.03, 0.03)   # 12 MB -> 6 sub-batches
    # every 7th read is replaced by random sequence: no anchors, no record -> the workers' in-place record slices have gaps
    # that the final compaction must close
    rng = np.random.default_rng(5)
    cat = cat.copy()
    for i in range(3, 3000, 7):
        lo, hi = int(roffs[i]), int(roffs[i + 1])
        cat[lo:hi] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, hi - lo)]
    r1 = c.map_batch(gi, cat, roffs)
    o = mm2.default_map_opts()
    o.want_stage_dump = 1                                              # forces the single-context path
    r2 = c.map_batch(gi, cat, roffs, o)
    assert r1.recs.size == r2.recs.size and 2000 < r1.recs.size < 3000 and (r1.recs == r2.recs).all()
    assert r1.stats["n_anchors"] == r2.stats["n_anchors"] and r1.stats["n_minimizers"] == r2.stats["n_minimizers"]
    oi = orc.Index.build(g, offs, ["p"], threads=8)
    names = ["n%d" % i for i in range(3000)]
    want, _ = oi.align_batch(cat, roffs, names, threads=8)
    assert r1.paf_lines(names) == want
    # batches larger than the resident limit go through the pipeline in pieces (here 4 MB pieces of 2 sub-batches each)
    os.environ["MM2_RESIDENT_MB"] = "4"
    try:
        r3 = c.map_batch(gi, cat, roffs)
    finally:
        del os.environ["MM2_RESIDENT_MB"]
    assert (r3.recs == r2.recs).all() and r3.stats["n_anchors"] == r2.stats["n_anchors"]
    # a batch that does not start at offset 0 of its buffer
    r4 = c.map_batch(gi, cat, roffs[1000:])
    want4 = [l for l in want if int(l.split("\t")[0][1:]) >= 1000]
    assert r4.recs.size == len(want4) and r4.paf_lines(names[1000:]) == want4
    c.close()


@pytest.mark.parametrize("world,k,b,w,hpc", [(2, 15, 14, 10, 0), (3, 19, 10, 10, 0), (8, 15, 14, 10, 0), (5, 17, 12, 5, 0), (4, 14, 12, 10, 0), (3, 15, 14, 10, 1)])
def test_multi_gpu_index_build_emulated_ranks(ctx, mm2, orc, gen, tmp_path, world, k, b, w, hpc):
    """bucket-sharded build with R virtual ranks on one GPU: byte-identical .mmi.
    Odd k / no HPC shards the sketch by TILES, i.e. inside sequences (every virtual rank only uploads its own bytes plus the
    halo; the rest of the sequence buffer is poisoned with N); even k and HPC shard by whole sequences."""
    g = gen.repeat_genome(61, 900_000, 0.3, 0.2)
    cuts = [0, 200_003, 200_004, 450_000, 450_017, 700_001, 900_000]
    seqs = [g[cuts[i]:cuts[i + 1]].tobytes() for i in range(len(cuts) - 1)]
    names = ["s%d" % i for i in range(len(seqs))]
    cat, offs = cases.cat_offs(seqs)
    gi = mm2.Index.build_sharded_emulated(ctx, world, cat, offs, names, w=w, k=k, b=b, flag=hpc)
    oi = orc.Index.build(cat, offs, names, w=w, k=k, b=b, flag=hpc, threads=8)
    assert gi.stats() == oi.stats() and gi.calc_mid_occ() == oi.calc_mid_occ()
    pg, po = str(tmp_path / "g.mmi"), str(tmp_path / "o.mmi")
    gi.save_to_mmi(pg)
    oi.save_mmi(po)
    assert open(pg, "rb").read() == open(po, "rb").read()
    if hpc:
        return
    # and it maps like the single-GPU index
    rc, ro = gen.reads(2, g[:200_000], np.array([0, 200_000], dtype=np.uint64), 20, 3000, 0.02, 0.02, 0.02)
    res = ctx.map_batch(gi, rc, ro, mm2.default_map_opts(w, k))
    want, _ = oi.align_batch(rc, ro, ["m%d" % i for i in range(20)], orc.AlignOpts.default(w, k), threads=4)
    assert res.paf_lines(["m%d" % i for i in range(20)]) == want


def test_multi_gpu_single_sequence_genome_is_split_inside_the_sequence(ctx, mm2, orc, gen, tmp_path):
    """one chromosome, 8 virtual ranks: every rank sketches 1/8 of it (the reference's rayon split cannot, index.rs:442-452)"""
    g = gen.genome(63, 3_000_000, 1e-3, 30.0)
    offs = np.array([0, g.size], dtype=np.uint64)
    plans = [mm2.shard_plan(offs, 10, 15, 0, 8, r) for r in range(8)]
    assert all(p["tile_path"] == 1 for p in plans) and plans[0]["lo"] == 0
    assert all(plans[r]["hi"] == plans[r + 1]["lo"] for r in range(7))
    assert max(p["upload_bytes"] for p in plans) < g.size // 8 + 5000
    gi = mm2.Index.build_sharded_emulated(ctx, 8, g, offs, ["chr1"])
    oi = orc.Index.build(g, offs, ["chr1"], threads=8)
    pg, po = str(tmp_path / "g.mmi"), str(tmp_path / "o.mmi")
    gi.save_to_mmi(pg)
    oi.save_mmi(po)
    assert open(pg, "rb").read() == open(po, "rb").read()


def test_two_contexts_on_two_devices_in_one_process(mm2, orc, gen, tmp_path):
    """per-device kernel attributes (dynamic shared memory opt-ins) and mm2_index_build_multi: needs >= 2 GPUs"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    g = gen.repeat_genome(65, 1_500_000, 0.3, 0.2)
    offs = np.array([0, g.size], dtype=np.uint64)
    cs = [mm2.Context(0), mm2.Context(1)]
    gis = mm2.Index.build_multi(cs, g, offs, ["m"])
    oi = orc.Index.build(g, offs, ["m"], threads=8)
    cat, roffs = gen.reads(3, g, offs, 50, 12000, 0.02, 0.02, 0.02)
    names = ["t%d" % i for i in range(50)]
    want, _ = oi.align_batch(cat, roffs, names, threads=8)
    os.environ["MM2_CHAIN_DENSE_MIN"] = "1"      # the CTA-per-read kernel needs the 95 KB opt-in on BOTH devices
    try:
        ds = [mm2.Context(0), mm2.Context(1)]
    finally:
        del os.environ["MM2_CHAIN_DENSE_MIN"]
    for c, gi in list(zip(cs, gis)) + list(zip(ds, gis)):
        assert gi.stats() == oi.stats()
        assert c.map_batch(gi, cat, roffs).paf_lines(names) == want
    p0, p1, po = (str(tmp_path / n) for n in ("0.mmi", "1.mmi", "o.mmi"))
    gis[0].save_to_mmi(p0)
    gis[1].save_to_mmi(p1)
    oi.save_mmi(po)
    assert open(p0, "rb").read() == open(po, "rb").read() == open(p1, "rb").read()


@pytest.mark.parametrize("min_cnt,min_score,extra", [(1, 10, {}), (1, 40, {}), (1, 5, dict(best_n=2, pri_ratio=0.5, mask_level=0.9)), (1, 10, dict(bw=100, bw_long=3000))])
def test_map_batch_general_path_single_anchor_chains(ctx, mm2, orc, gen, min_cnt, min_score, extra):
    """-n 1: chain_dp_all returns many single-anchor chains (F3) -> rescue / merge / select / several PAF lines per read"""
    g = gen.repeat_genome(91, 800_000, 0.3, 0.2)
    offs = np.array([0, g.size], dtype=np.uint64)
    gi = mm2.Index.build(ctx, g, offs, ["gp"])
    oi = orc.Index.build(g, offs, ["gp"], threads=8)
    cat, roffs = gen.reads(6, g, offs, 40, 2500, 0.03, 0.03, 0.03)
    names = ["g%d" % i for i in range(40)]
    o = mm2.default_map_opts()
    oo = orc.AlignOpts.default()
    o.min_cnt = oo.min_cnt = min_cnt
    o.min_chain_score = oo.min_chain_score = min_score
    for kk, vv in extra.items():
        setattr(o, kk, vv)
        setattr(oo, kk, vv)
    res = ctx.map_batch(gi, cat, roffs, o)
    want, st = oi.align_batch(cat, roffs, names, oo, threads=8)
    got = res.paf_lines(names)
    if min_score <= 15:
        assert len(want) > 40                  # really several lines per read
    assert got == want
    assert res.stats["n_rescued"] == st.n_rescued


def test_map_batch_query_wk_differs_from_index(ctx, mm2, orc, gen):
    """align sketches the query with the CLI's w/k, dv uses the index's (paf.rs:156)"""
    g = gen.genome(93, 600_000)
    offs = np.array([0, g.size], dtype=np.uint64)
    gi = mm2.Index.build(ctx, g, offs, ["wk"], w=10, k=15)
    oi = orc.Index.build(g, offs, ["wk"], w=10, k=15, threads=8)
    cat, roffs = gen.reads(7, g, offs, 30, 3000, 0.01, 0.01, 0.01)
    names = ["k%d" % i for i in range(30)]
    for w, k in ((5, 15), (10, 14), (12, 15)):
        res = ctx.map_batch(gi, cat, roffs, mm2.default_map_opts(w, k))
        want, _ = oi.align_batch(cat, roffs, names, orc.AlignOpts.default(w, k), threads=8)
        assert res.paf_lines(names) == want, (w, k)


def test_map_batch_splits_when_anchors_exceed_memory(mm2, orc, gen):
    """a batch whose anchors + DP state do not fit is mapped in halves (recursively); records must not change"""
    g = gen.repeat_genome(97, 1_000_000, 0.4, 0.2)
    offs = np.array([0, g.size], dtype=np.uint64)
    c = mm2.Context(0)
    gi = mm2.Index.build(c, g, offs, ["m"])
    cat, roffs = gen.reads(5, g, offs, 64, 4000, 0.03, 0.03, 0.03)
    d_cat = None
    import torch
    d_cat = torch.from_numpy(cat.copy()).cuda()
    d_off = torch.from_numpy(roffs.astype(np.int64)).cuda()
    r_full = c.map_batch(gi, None, roffs, device_ptrs=(d_cat.data_ptr(), d_off.data_ptr()))
    os.environ["MM2_ANCHOR_BUDGET_MB"] = "1"
    try:
        c2 = mm2.Context(0)      # fresh arenas: everything has to "grow", and 1 MB never suffices for more than a few reads
        r_split = c2.map_batch(gi, None, roffs, device_ptrs=(d_cat.data_ptr(), d_off.data_ptr()))
    finally:
        del os.environ["MM2_ANCHOR_BUDGET_MB"]
    assert r_full.recs.size == r_split.recs.size and (r_full.recs == r_split.recs).all()
    assert r_full.stats == r_split.stats
    names = ["s%d" % i for i in range(64)]
    oi = orc.Index.build(g, offs, ["m"], threads=8)
    want, _ = oi.align_batch(cat, roffs, names, threads=8)
    assert r_split.paf_lines(names) == want
    c.close(); c2.close()


def test_fuzz_sketch_parameters(ctx, orc):
    """seeded differential fuzz of the sketch kernels over (w, k, alphabet, N density, length)"""
    rng = np.random.default_rng(2024)
    for it in range(120):
        k = int(rng.integers(1, 29))
        w = int(rng.choice([1, 2, 3, 5, 8, 9, 10, 11, 16, 19, 33, 64, 100, 255]))
        n = int(rng.choice([1, 7, 30, 100, 500, 2030, 2050, 4100, 9000]))
        alpha = rng.choice(["ACGT", "ACGTN", "AC", "ACGTacgtRYn", "AT"])
        s = cases.rnd_seq(rng, n, alpha)
        if rng.random() < 0.3 and n > 50:
            b = bytearray(s)
            p = int(rng.integers(0, n - 20))
            b[p:p + 20] = b[max(0, p - 20):p][:20].ljust(20, b"A")   # local duplication -> equal keys inside a window
            s = bytes(b)
        hpc = bool(rng.random() < 0.15)
        a, b_ = ctx.sketch_sequence(s, w, k, rid=int(rng.integers(0, 5)) * 2, is_hpc=hpc), None
        b_ = orc.sketch(s, w, k, rid=int(a["rid_pos_strand"][0] >> np.uint64(32)) if a.size else 0, is_hpc=hpc)
        _eq(a, b_, "fuzz %d: w=%d k=%d n=%d alpha=%s hpc=%s" % (it, w, k, n, alpha, hpc))


def test_fuzz_map_options(ctx, mm2, orc, gen):
    """seeded differential fuzz of mm2_map_batch over the align flag surface"""
    rng = np.random.default_rng(7)
    g = gen.repeat_genome(123, 700_000, 0.3, 0.2)
    seqs = [g[:400_000].tobytes(), b"NN", g[400_000:].tobytes()]
    cat, offs = cases.cat_offs(seqs)
    names = ["a", "pad", "b"]
    for k, w in ((15, 10), (17, 7)):
        gi = mm2.Index.build(ctx, cat, offs, names, w=w, k=k, b=12)
        oi = orc.Index.build(cat, offs, names, w=w, k=k, b=12, threads=8)
        for it in range(6):
            n = 25
            ln = int(rng.choice([300, 1500, 4000]))
            e = float(rng.choice([0.0, 0.01, 0.05]))
            rc, ro = gen.reads(int(rng.integers(1, 1000)), cat, offs, n, ln, e, e, e)
            o, oo = mm2.default_map_opts(w, k), orc.AlignOpts.default(w, k)
            for name, val in (("max_gap", int(rng.choice([500, 5000, 20000]))), ("min_cnt", int(rng.choice([1, 2, 3, 6]))),
                              ("min_chain_score", int(rng.choice([5, 40, 200]))), ("bw", int(rng.choice([-1, 50, 500, 3000]))),
                              ("bw_long", int(rng.choice([-1, 1000, 20000]))), ("frac_top_repetitive", float(rng.choice([2e-4, 1e-2, 0.3]))),
                              ("best_n", int(rng.choice([0, 1, 5]))), ("pri_ratio", float(rng.choice([0.3, 0.8]))),
                              ("mask_level", float(rng.choice([0.1, 0.5, 0.95])))):
                setattr(o, name, val)
                setattr(oo, name, val)
            if o.bw < 0:
                o.bw_long = oo.bw_long = -1       # `-r` absent: neither value is given (main.rs:202-208)
            qn = ["f%d" % i for i in range(n)]
            res = ctx.map_batch(gi, rc, ro, o)
            want, _ = oi.align_batch(rc, ro, qn, oo, threads=8)
            assert res.paf_lines(qn) == want, "fuzz k=%d it=%d opts=%s" % (k, it, {f: getattr(o, f) for f, _ in o._fields_})


@pytest.mark.parametrize("w,k", [(10, 14), (10, 16), (5, 20), (11, 28)])
def test_map_batch_even_k_matching_index(ctx, mm2, orc, gen, w, k):
    """even k with the SAME w/k on the index and the query: the default (closed-form dv) record path, not the w/k-mismatch
    tail.  Palindromic k-mers are skipped by the sketch (sketch.rs:67), so the literal sketch kernel, the lookup and the
    rank-based dv (paf.rs:174-191) all see them."""
    g = gen.repeat_genome(131, 900_000, 0.3, 0.2) if k < 20 else gen.genome(132, 900_000, 1e-3, 30.0)
    # palindrome-rich inserts: (AT)n / ACGT-repeat blocks make many k-mers equal to their reverse complement
    g = g.copy()
    for p in range(10_000, 800_000, 90_000):
        g[p:p + 300] = np.frombuffer((b"ACGT" * 75), dtype=np.uint8)
        g[p + 1000:p + 1200] = np.frombuffer((b"AT" * 100), dtype=np.uint8)
    offs = np.array([0, g.size], dtype=np.uint64)
    gi = mm2.Index.build(ctx, g, offs, ["ev"], w=w, k=k)
    oi = orc.Index.build(g, offs, ["ev"], w=w, k=k, threads=8)
    assert gi.stats() == oi.stats()
    cat, roffs = gen.reads(k, g, offs, 60, 3000, 0.01, 0.01, 0.01)
    reads = [cat[int(roffs[i]):int(roffs[i + 1])].tobytes() for i in range(60)]
    reads += [g[9_500:12_000].tobytes(), g[99_800:101_500].tobytes()]          # reads across the palindromic blocks
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    reads += [g[189_000:192_000].tobytes().translate(comp)[::-1]]
    cat2, ro2 = cases.cat_offs(reads)
    names = ["e%d" % i for i in range(len(reads))]
    res, st = _map_compare(ctx, mm2, orc, gi, oi, cat2, ro2, names, (w, k))
    assert res.recs.size >= 55
    # and without the stage dump (the pipelined / in-place record path)
    _map_compare(ctx, mm2, orc, gi, oi, cat2, ro2, names, (w, k), dump=False)


def test_config1_full_size_mmi_and_600bp_read(ctx, mm2, orc, gen, tmp_path):
    """BASELINE configs[0] at its stated size (SURVEY.md §8d C1 stand-in): one 145,138,636-bp sequence `chr8`, index k=15 w=10;
    the .mmi written by the GPU build must be byte-identical (sha256 over the whole file) to the CPU oracle's, and the 600-bp
    read genome[5,999,340 .. 5,999,940) with 4 substitutions must give the same PAF line."""
    import hashlib
    L = 145_138_636
    g = gen.genome(0xB2000002, L)
    offs = np.array([0, L], dtype=np.uint64)
    gi = mm2.Index.build(ctx, g, offs, ["chr8"], w=10, k=15, b=14)
    oi = orc.Index.build(g, offs, ["chr8"], w=10, k=15, b=14, threads=os.cpu_count() or 8)
    assert gi.stats() == oi.stats()
    assert gi.calc_mid_occ(2e-4) == oi.calc_mid_occ(2e-4)

    def sha(path):
        h = hashlib.sha256()
        with open(path, "rb") as f:
            while True:
                blk = f.read(1 << 24)
                if not blk:
                    break
                h.update(blk)
        return h.hexdigest(), os.path.getsize(path)

    pg, po = str(tmp_path / "g.mmi"), str(tmp_path / "o.mmi")
    gi.save_to_mmi(pg)
    hg = sha(pg)
    os.remove(pg)
    oi.save_mmi(po)
    ho = sha(po)
    os.remove(po)
    assert hg == ho, (hg, ho)
    q = bytearray(g[5_999_340:5_999_940].tobytes())
    for p in (57, 211, 388, 540):
        q[p] = ord("ACGT"[("ACGT".index(chr(q[p])) + 1) % 4])
    qc, qo = cases.cat_offs([bytes(q)])
    res = ctx.map_batch(gi, qc, qo)
    want, _ = oi.align_batch(qc, qo, ["read600"])
    got = res.paf_lines(["read600"])
    assert got == want and len(got) == 1
    f = got[0].split("\t")
    assert f[5] == "chr8" and int(f[6]) == L and abs(int(f[7]) - 5_999_340) < 40 and f[4] == "+"
    # a sample of BASELINE configs[1] reads on the same full-size index, every stage compared
    cat, roffs = gen.reads(0xB2001002, g, offs, 200, 10_000, 0.0333, 0.0333, 0.0333)
    _map_compare(ctx, mm2, orc, gi, oi, cat, roffs, ["r%06d" % i for i in range(200)])


def test_mmi_foreign_layouts_no_seq_any_order_any_b(ctx, mm2, orc, gen, tmp_path):
    """indexes written by other tools — entries of a bucket in any order (C minimap2 dumps khash slot
    order, the reference writes random HashMap order), MM_I_NO_SEQ files without the packed sequence, b != 14."""
    import mmi_util
    g = gen.repeat_genome(141, 700_000, 0.3, 0.2)
    seqs = [g[:400_000].tobytes(), b"N", g[400_000:].tobytes()]
    cat, offs = cases.cat_offs(seqs)
    names = ["ca", "pad", "cb"]
    rc, ro = gen.reads(5, cat, offs, 40, 3000, 0.02, 0.02, 0.02)
    qn = ["f%d" % i for i in range(40)]
    rng = np.random.default_rng(3)
    for b in (10, 14, 17):
        oi = orc.Index.build(cat, offs, names, w=10, k=15, b=b, threads=8)
        want, _ = oi.align_batch(rc, ro, qn, threads=8)
        po = str(tmp_path / ("o%d.mmi" % b))
        oi.save_mmi(po)
        canon = open(po, "rb").read()
        m = mmi_util.parse(canon)
        assert mmi_util.serialise(m) == canon                       # the fixture writer reproduces the layout
        # (1) every bucket's entries in random order
        m["buckets"] = [(p, ent[rng.permutation(len(ent))]) for p, ent in m["buckets"]]
        ps = str(tmp_path / ("s%d.mmi" % b))
        open(ps, "wb").write(mmi_util.serialise(m))
        gs = mm2.Index.load_from_mmi(ctx, ps)
        assert gs.stats() == oi.stats() and gs.calc_mid_occ() == oi.calc_mid_occ() and gs.b == b
        assert ctx.map_batch(gs, rc, ro).paf_lines(qn) == want
        pb = str(tmp_path / ("b%d.mmi" % b))
        gs.save_to_mmi(pb)
        assert open(pb, "rb").read() == canon                       # order-independent: written back canonically
        # (2) MM_I_NO_SEQ: flag bit 1 set, no packed sequence at the end of the file
        m["flag"] |= 2
        pn = str(tmp_path / ("n%d.mmi" % b))
        open(pn, "wb").write(mmi_util.serialise(m, with_seq=False))
        gn = mm2.Index.load_from_mmi(ctx, pn)
        assert gn.stats() == oi.stats() and (gn.flag & 2)
        assert ctx.map_batch(gn, rc, ro).paf_lines(qn) == want      # mapping never touches S
        with pytest.raises(mm2.Mm2Error) as e:
            gn.get_ref_subseq(0, 0, 10)
        assert e.value.code == mm2.MM2_E_FORMAT
        pn2 = str(tmp_path / ("n2_%d.mmi" % b))
        gn.save_to_mmi(pn2)
        m2 = mmi_util.parse(open(pn2, "rb").read())
        assert m2["flag"] & 2 and len(m2["S"]) == 0
        # a NO_SEQ flag on a file that is simply truncated elsewhere is still an error
        with pytest.raises(mm2.Mm2Error):
            open(pn, "wb").write(mmi_util.serialise(m, with_seq=False)[:-9])
            mm2.Index.load_from_mmi(ctx, pn)
        for x in (gs, gn):
            x.close()


def test_mmi_khash_slot_order_writer(ctx, mm2, orc, gen, tmp_path):
    """.mmi with every bucket in the slot order of C minimap2's khash table.  Checked against an independent
    Python model of klib's khash (tests/mmi_util.py), for bucket sizes below and above the 0.77 load factor of the initial
    table (the latter force khash's in-place rehash with its kick-out chain)."""
    import mmi_util
    g = gen.repeat_genome(143, 600_000, 0.3, 0.2)
    offs = np.array([0, g.size], dtype=np.uint64)
    for b in (6, 12):     # b = 6: ~1500 keys per bucket; b = 12: ~25
        gi = mm2.Index.build(ctx, g, offs, ["k"], w=10, k=15, b=b)
        pc, pk = str(tmp_path / "c.mmi"), str(tmp_path / "k.mmi")
        gi.save_to_mmi(pc)
        gi.save_to_mmi_khash(pk)
        mc, mk = mmi_util.parse(open(pc, "rb").read()), mmi_util.parse(open(pk, "rb").read())
        assert mc["seqs"] == mk["seqs"] and mc["S"] == mk["S"]
        n_resized = 0
        for (p1, e1), (p2, e2) in zip(mc["buckets"], mk["buckets"]):
            assert (p1 == p2).all() and len(e1) == len(e2)
            want = mmi_util.khash_order([(int(a), int(c)) for a, c in e1])
            assert [(int(a), int(c)) for a, c in e2] == want
            nb0 = max(4, mmi_util.Khash._roundup32(len(e1))) if len(e1) else 0
            n_resized += int(len(e1) > int(nb0 * 0.77 + 0.5))
        assert n_resized > 0                                        # the rehash path was exercised
        g2 = mm2.Index.load_from_mmi(ctx, pk)
        pr = str(tmp_path / "r.mmi")
        g2.save_to_mmi(pr)
        assert open(pr, "rb").read() == open(pc, "rb").read()
        gi.close(); g2.close()


def test_cli_streaming_ingest_bounded_batches(ctx, mm2, orc, gen, tmp_path):
    """`align --all-reads` streams the query file in batches (a reader thread parses batch i + 1 into
    page-locked memory while the GPU maps batch i); with --batch-mb 1 a 9 MB file goes through in ~9 batches and the PAF must
    be the same lines in the same (input) order; FASTA with wrapped lines and FASTQ; reads without a hit leave no line."""
    import subprocess
    exe = os.path.join(os.path.dirname(mm2.LIB_PATH), "mm2rs")
    g = gen.genome(151, 1_500_000)
    offs = np.array([0, g.size], dtype=np.uint64)
    oi = orc.Index.build(g, offs, ["ref"], threads=8)
    mmi = str(tmp_path / "ref.mmi")
    oi.save_mmi(mmi)
    n = 2300
    rc, ro = gen.reads(9, g, offs, n, 4000, 0.02, 0.02, 0.02)
    rc = rc.copy()
    rng = np.random.default_rng(4)
    for i in range(5, n, 11):                       # unrelated reads: no anchors, no PAF line
        rc[int(ro[i]):int(ro[i + 1])] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 4000)]
    names = ["s%05d" % i for i in range(n)]
    want, _ = oi.align_batch(rc, ro, names, threads=8)
    fa, fq = str(tmp_path / "q.fa"), str(tmp_path / "q.fq")
    with open(fa, "wb") as f, open(fq, "wb") as f2:
        for i in range(n):
            sq = rc[int(ro[i]):int(ro[i + 1])].tobytes()
            f.write(b">" + names[i].encode() + b" desc\n")
            for j in range(0, len(sq), 80):
                f.write(sq[j:j + 80] + (b"\r\n" if i % 3 == 0 else b"\n"))
            f2.write(b"@" + names[i].encode() + b"\n" + sq + b"\n+\n" + b"@" * len(sq) + b"\n")
    for path in (fa, fq):
        for mb in ("1", "3", "1024"):
            out = subprocess.run([exe, "align", mmi, path, "--all-reads", "--batch-mb", mb], capture_output=True, text=True)
            assert out.returncode == 0, out.stderr
            assert out.stdout == "\n".join(want) + "\n", (path, mb)
    # without --all-reads: the first record only (main.rs:92-103)
    out = subprocess.run([exe, "align", mmi, fa], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout == want[0] + "\n"
    # khash-order dump from the CLI loads back and maps the same
    fref, mk = str(tmp_path / "ref.fa"), str(tmp_path / "k.mmi")
    with open(fref, "wb") as f:
        f.write(b">ref\n" + g.tobytes() + b"\n")
    out = subprocess.run([exe, "index", "-d", mk, "--khash-order", fref], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert open(mk, "rb").read() != open(mmi, "rb").read() and os.path.getsize(mk) == os.path.getsize(mmi)
    out = subprocess.run([exe, "align", mk, fa, "--all-reads"], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout == "\n".join(want) + "\n"
    import torch
    if torch.cuda.device_count() >= 2:              # extension: `index --gpus 2` = mm2_index_build_multi
        m2 = str(tmp_path / "m2.mmi")
        out = subprocess.run([exe, "index", "-d", m2, "--gpus", "2", fref], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        assert open(m2, "rb").read() == open(mmi, "rb").read()
