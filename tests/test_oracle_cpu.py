"""CPU suite: the oracle against itself / brute force / the numpy models of the GPU reformulations.
(The reference ships no golden vectors — SURVEY.md F2 — so these are property tests; parity stays "unpinned".)"""
import os

import numpy as np
import pytest

import cases
import models


def test_sketch_model_matches_oracle(orc):
    for w, k in [(10, 15), (10, 19), (3, 5), (1, 7)]:
        for name, s in cases.sketch_cases():
            if len(s) > 3100:
                continue
            a = orc.sketch(s, w, k, rid=2)
            m = models.sketch_model(s, w, k, rid=2)
            assert a.size == m.size and (a == m).all(), (name, w, k)


def test_sketch_model_v4_matches_oracle(orc):
    """the by-position restatement of sketch_tile_kernel_v4 (opening of the key sequence + by-step marking of dirty
    chunks) on small regions, so that tile boundaries, N runs, sequence ends and tie-rich windows all meet"""
    for w, k in [(10, 15), (9, 5), (19, 7), (12, 3)]:
        for name, s in cases.sketch_cases():
            if len(s) > 3100:
                continue
            a = orc.sketch(s, w, k, rid=2)
            m = models.sketch_model_v4(s, w, k, rid=2, region=256 if w < 20 else 512)
            assert a.size == m.size and (a == m).all(), (name, w, k)
    rng = np.random.default_rng(12)
    for it in range(400):
        w = int(rng.integers(9, 24))
        k = int(rng.choice([3, 5, 7, 9, 15]))
        s = cases.rnd_seq(rng, int(rng.integers(1, 700)), ["AC", "ACGT", "A", "ACGTN", "AT", "ACGTNNN"][int(rng.integers(0, 6))])
        a = orc.sketch(s, w, k, rid=1)
        m = models.sketch_model_v4(s, w, k, rid=1, region=int(rng.choice([128, 256])))
        assert a.size == m.size and (a == m).all(), (it, w, k, len(s))
    # -H: the span (sum of the remaining homopolymer runs of the last k bases) is part of the key; k-mers whose span reaches 256
    # drop out without resetting l
    for it in range(300):
        w = int(rng.integers(9, 24))
        k = int(rng.choice([3, 5, 9, 15, 19]))
        s = cases.rnd_seq(rng, int(rng.integers(1, 900)), ["AC", "ACGT", "AAAAC", "AACCGGTTN", "A" * 40 + "C"][int(rng.integers(0, 5))])
        if it % 5 == 0:
            s = s[:len(s) // 3] + b"G" * int(rng.integers(200, 400)) + s[len(s) // 3:]
        a = orc.sketch(s, w, k, rid=1, is_hpc=True)
        m = models.sketch_model_v4(s, w, k, rid=1, region=int(rng.choice([128, 256])), is_hpc=True)
        assert a.size == m.size and (a == m).all(), ("hpc", it, w, k, len(s))


def test_sketch_properties(orc):
    rng = np.random.default_rng(5)
    s = cases.rnd_seq(rng, 50_000)
    mv = orc.sketch(s, 10, 15)
    pos = (mv["rid_pos_strand"] >> np.uint64(1)) & np.uint64(0xffffffff)
    assert (np.diff(pos.astype(np.int64)) > 0).all()          # strictly position-increasing, no duplicates
    assert 0.17 < mv.size / len(s) < 0.20                     # ~2/(w+1)
    assert ((mv["key_span"] & np.uint64(0xff)) == 15).all()
    # reverse complement gives the same keys (canonical k-mers), opposite strands
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    rc = s.translate(comp)[::-1]
    mv2 = orc.sketch(rc, 10, 15)
    assert set(mv["key_span"].tolist()) == set(mv2["key_span"].tolist()) or abs(mv.size - mv2.size) < 20


def test_short_sequences_emit_one(orc):
    rng = np.random.default_rng(6)
    for n in range(15, 24):
        mv = orc.sketch(cases.rnd_seq(rng, n), 10, 15)
        assert mv.size == 1   # sketch.rs:99 end-of-sequence emit, Appendix A item 7
    assert orc.sketch(cases.rnd_seq(rng, 14), 10, 15).size == 0


def test_index_get_vs_bruteforce_and_mmi_roundtrip(orc, gen, tmp_path):
    g = gen.repeat_genome(3, 200_000, 0.3, 0.2)
    seqs = [bytes(g[:120_000]), b"N", bytes(g[120_000:])]
    cat, offs = cases.cat_offs(seqs)
    idx = orc.Index.build(cat, offs, ["a", "n", "b"], w=10, k=15, b=10, threads=4)
    allm = np.concatenate([orc.sketch(s, 10, 15, rid=i) for i, s in enumerate(seqs)])
    keys = allm["key_span"] >> np.uint64(8)
    uk, cnt = np.unique(keys, return_counts=True)
    nk, avg_occ, spacing, tl = idx.stats()
    assert nk == uk.size and tl == len(g) + 1
    assert abs(avg_occ - allm.size / uk.size) < 1e-9
    rng = np.random.default_rng(0)
    for j in rng.choice(uk.size, 300, replace=False):
        kind, occ = idx.get(int(uk[j]))
        want = np.sort(allm["rid_pos_strand"][keys == uk[j]])
        assert kind == (1 if cnt[j] == 1 else 2)
        assert (occ == want).all()
    assert idx.get(int(uk.max()) + 1)[0] == 0
    # calc_mid_occ: index.rs:124-141
    srt = np.sort(cnt)
    want = int(srt[min(int((1.0 - float(np.float32(2e-4))) * srt.size), srt.size - 1)]) + 1
    assert idx.calc_mid_occ(2e-4) == want
    # .mmi round trip through the oracle's own loader: identical bytes when written again (canonical order)
    p1, p2, p3 = (str(tmp_path / n) for n in ("a.mmi", "b.mmi", "c.idx"))
    idx.save_mmi(p1)
    idx2 = orc.Index.load_mmi(p1)
    idx2.save_mmi(p2)
    assert open(p1, "rb").read() == open(p2, "rb").read()
    assert idx2.stats() == idx.stats() and idx2.params() == idx.params()
    idx.save_native(p3)
    idx3 = orc.Index.load_native(p3)
    assert idx3.stats() == idx.stats()
    hdr = open(p1, "rb").read(24)
    assert hdr[:4] == b"MMI\x02" and np.frombuffer(hdr[4:24], dtype="<u4").tolist() == [10, 15, 10, 3, 0]


def test_chain_model_matches_oracle(orc, gen):
    g = gen.repeat_genome(77, 200_000, 0.4, 0.2)
    offs = np.array([0, g.size], dtype=np.uint64)
    idx = orc.Index.build(g, offs, ["c"], threads=4)
    cat, roffs = gen.reads(5, g, offs, 6, 2000, 0.02, 0.02, 0.02)
    for r in range(6):
        q = cat[int(roffs[r]):int(roffs[r + 1])]
        mv = orc.filter_query_minimizers(orc.sketch(q, 10, 15))
        a = idx.anchors(mv, q.size, 60)
        if a.size > 4000:
            a = a[:4000]
        for bw, skip in ((500, 25), (20000, 3)):
            p = orc.default_chain_params(15)
            p.bw = bw
            p.max_chain_skip = skip
            o = orc.chain_dp_all(a, p)
            f, pp, v = models.chain_fwd_model(a, p)
            assert (f == o["f"]).all() and (pp == o["pprev"]).all() and (v == o["v"]).all()


def test_default_params_give_single_fallback_chain(orc, gen):
    """with -n >= 2 chain_dp_all returns exactly one chain, the fallback (last argmax of f, score v)."""
    g = gen.genome(11, 300_000)
    offs = np.array([0, g.size], dtype=np.uint64)
    idx = orc.Index.build(g, offs, ["c"], threads=4)
    cat, roffs = gen.reads(9, g, offs, 5, 5000, 0.03, 0.03, 0.03)
    for r in range(5):
        q = cat[int(roffs[r]):int(roffs[r + 1])]
        a = idx.anchors(orc.filter_query_minimizers(orc.sketch(q, 10, 15)), q.size, 10)
        o = orc.chain_dp_all(a, orc.default_chain_params(15))
        assert len(o["chains"]) == 1
        f = o["f"]
        best = int(np.nonzero(f == f.max())[0][-1])
        ch = o["chains"][0]
        assert ch[-1] == best and o["scores"][0] == o["v"][best]
        walk = []
        i = best
        while i >= 0:
            walk.append(i)
            i = int(o["pprev"][i])
        assert walk[::-1] == ch.tolist()
        p = orc.default_chain_params(15)
        p.min_cnt, p.min_chain_score = 1, 10
        o2 = orc.chain_dp_all(a, p)
        assert all(len(c) == 1 for c in o2["chains"])


def test_odd_rid_anchor_sign_extension(orc):
    """anchors to an odd rid carry 0xFFFFFFFF in the rid/strand bits; the PAF stage then 'panics'."""
    rng = np.random.default_rng(3)
    s0, s1 = cases.rnd_seq(rng, 3000), cases.rnd_seq(rng, 3000)
    cat, offs = cases.cat_offs([s0, s1])
    idx = orc.Index.build(cat, offs, ["a", "b"])
    q = s1[500:1500]
    a = idx.anchors(orc.sketch(q, 10, 15), len(q), 10)
    assert a.size > 0 and ((a["x"] >> np.uint64(32)) & np.uint64(0x7fffffff) == 0x7fffffff).all()
    lines, st = idx.align_batch(np.frombuffer(q, dtype=np.uint8), np.array([0, len(q)], dtype=np.uint64), ["q"])
    assert lines == [] and st.n_panic == 1


def test_paf_line_shape(orc, gen):
    g = gen.genome(0xB2000001, 400_000)
    offs = np.array([0, g.size], dtype=np.uint64)
    idx = orc.Index.build(g, offs, ["chr8"], threads=4)
    read = bytearray(g[59_340:59_940].tobytes())
    for p in (100, 250, 400, 555):
        read[p] = ord("A") if read[p] != ord("A") else ord("C")
    lines, st = idx.align_batch(np.frombuffer(bytes(read), dtype=np.uint8), np.array([0, 600], dtype=np.uint64), ["read1"])
    assert len(lines) == 1
    f = lines[0].split("\t")
    assert f[0] == "read1" and f[1] == "600" and f[4] == "+" and f[5] == "chr8" and f[6] == "400000" and f[11] == "60"
    assert f[12] == "tp:A:P" and f[15] == "s2:i:0" and f[17] == "rl:i:0" and len(f[16].split(":")[2].split(".")[1]) == 4
    assert abs(int(f[7]) - 59_340) < 30 and abs(int(f[8]) - 59_940) < 30


# ---- committed golden fixtures (tests/golden/make_golden.py) -----------------------------------------------------
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_oracle_matches_golden_fixtures(orc, gen):
    import hashlib
    d = np.load(os.path.join(GOLD, "sketch.npz"))
    for i in range(int(d["n"])):
        mv = orc.sketch(d["seq%d" % i].tobytes(), int(d["w"][i]), int(d["k"][i]), rid=1)
        assert (mv["key_span"] == d["key%d" % i]).all() and (mv["rid_pos_strand"] == d["val%d" % i]).all(), i
    g = gen.genome(0xB2000001, 200_000)
    offs = np.array([0, g.size], dtype=np.uint64)
    idx = orc.Index.build(g, offs, ["chr8"], threads=4)
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "x.mmi")
        idx.save_mmi(p)
        sha, size = open(os.path.join(GOLD, "index_200k.sha256")).read().split()
        assert hashlib.sha256(open(p, "rb").read()).hexdigest() == sha and os.path.getsize(p) == int(size)
    cat, roffs = gen.reads(0xB2001001, g, offs, 24, 3000, 0.02, 0.02, 0.02)
    lines, _ = idx.align_batch(cat, roffs, ["r%02d" % i for i in range(24)])
    assert lines == open(os.path.join(GOLD, "align_24reads.paf")).read().split("\n")[:-1]
    s = np.load(os.path.join(GOLD, "read0_stages.npz"))
    a = idx.anchors(orc.filter_query_minimizers(orc.sketch(cat[:3000], 10, 15)), 3000, max(10, idx.calc_mid_occ()))
    assert (a["x"] == s["ax"]).all() and (a["y"] == s["ay"]).all()
    o = orc.chain_dp_all(a, orc.default_chain_params(15))
    assert (o["f"] == s["f"]).all() and (o["pprev"] == s["pprev"]).all() and (o["chains"][0] == s["chain0"]).all()


@pytest.mark.parametrize("density,max_skip,max_iter,bw", [(300, 25, 5000, 500), (300, 3, 200, 500), (120, 25, 100, 100000), (600, 0, 5000, 500),
                                                         (40, 25, 33, 500), (300, 1000, 700, 500)])
def test_chain_dense_tile_model_matches_oracle(orc, density, max_skip, max_iter, bw):
    """the tile pipeline of chain_dense_kernel (summaries of the far window + sequential walk) against the oracle:
    f, pprev, v and the inner-loop cell count"""
    rng = np.random.default_rng(density + max_skip + max_iter)
    n = 1300
    parts = []
    for rid, rev, cnt in ((0, 0, n // 2), (0, 1, n // 4), (2, 0, n // 4)):
        extent = max(cnt * 5000 // density, 50)
        rpos = np.sort(rng.integers(20, 20 + extent, cnt)).astype(np.int64)
        diag = rng.choice([0, 0, 0, 7, -13, 400, -2500], cnt).astype(np.int64)
        run = rng.integers(0, 4, cnt) == 0
        diag[1:][run[1:]] = diag[:-1][run[1:]]
        qpos = np.clip(rpos + diag + rng.integers(-3, 4, cnt) * (rng.integers(0, 3, cnt) == 0), 14, None)
        x = (np.uint64(rev) << np.uint64(63)) | (np.uint64(rid) << np.uint64(32)) | rpos.astype(np.uint64)
        y = (np.uint64(15) << np.uint64(32)) | qpos.astype(np.uint64)
        parts.append(np.stack([x, y], axis=1))
    xy = np.concatenate(parts)
    xy = xy[np.lexsort((xy[:, 1], xy[:, 0]))]
    a = np.zeros(xy.shape[0], dtype=orc.ANCHOR_DT)
    a["x"], a["y"] = xy[:, 0], xy[:, 1]
    p = orc.default_chain_params(15)
    p.max_chain_skip, p.max_chain_iter, p.bw = max_skip, max_iter, bw
    o = orc.chain_dp_all(a, p)
    f, pp, v, cells = models.chain_dense_model(a, p)
    assert (f == o["f"]).all() and (pp == o["pprev"]).all() and (v == o["v"]).all()
    assert cells == o["cells"]
