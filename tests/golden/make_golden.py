"""Generates tests/golden/*.npz|.paf from the CPU oracle (the reference itself cannot be built here — no rustc — and
ships no vectors of its own).  These fixtures pin the ORACLE against regressions and give the GPU tests a
vector set that does not depend on recomputing the oracle;  they are not an independent pin of the reference.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import orc  # noqa: E402
from tools import gen  # noqa: E402
import cases  # noqa: E402


def main():
    # 1. sketch vectors: (sequence, w, k) -> minimizer arrays
    seqs, ws, ks, outs = [], [], [], []
    for name, s in cases.sketch_cases(seed=1):
        if len(s) > 6500:
            continue
        for w, k in ((10, 15), (10, 19), (5, 4)):
            seqs.append(np.frombuffer(s, dtype=np.uint8))
            ws.append(w); ks.append(k)
            outs.append(orc.sketch(s, w, k, rid=1))
    np.savez_compressed(os.path.join(HERE, "sketch.npz"), n=len(seqs), w=np.array(ws), k=np.array(ks),
                        **{"seq%d" % i: s for i, s in enumerate(seqs)},
                        **{"key%d" % i: o["key_span"] for i, o in enumerate(outs)},
                        **{"val%d" % i: o["rid_pos_strand"] for i, o in enumerate(outs)})
    # 2. a small genome (seeded generator) + reads -> .mmi checksum, anchors/DP of one read, PAF lines
    g = gen.genome(0xB2000001, 200_000)
    offs = np.array([0, g.size], dtype=np.uint64)
    idx = orc.Index.build(g, offs, ["chr8"], threads=4)
    tmp = os.path.join(HERE, "_tmp.mmi")
    idx.save_mmi(tmp)
    import hashlib
    sha = hashlib.sha256(open(tmp, "rb").read()).hexdigest()
    size = os.path.getsize(tmp)
    os.remove(tmp)
    cat, roffs = gen.reads(0xB2001001, g, offs, 24, 3000, 0.02, 0.02, 0.02)
    names = ["r%02d" % i for i in range(24)]
    lines, st = idx.align_batch(cat, roffs, names)
    open(os.path.join(HERE, "align_24reads.paf"), "w").write("\n".join(lines) + "\n")
    q = cat[:3000]
    mv = orc.filter_query_minimizers(orc.sketch(q, 10, 15))
    a = idx.anchors(mv, 3000, max(10, idx.calc_mid_occ()))
    o = orc.chain_dp_all(a, orc.default_chain_params(15))
    np.savez_compressed(os.path.join(HERE, "read0_stages.npz"), ax=a["x"], ay=a["y"], f=o["f"], v=o["v"], pprev=o["pprev"],
                        chain0=o["chains"][0], score0=o["scores"][0], stats=np.array(idx.stats()[:1] + idx.stats()[3:], dtype=np.uint64),
                        mid_occ=idx.calc_mid_occ(), mmi_size=size)
    open(os.path.join(HERE, "index_200k.sha256"), "w").write("%s  %d\n" % (sha, size))
    print("golden written:", len(seqs), "sketch vectors,", len(lines), "PAF lines, mmi sha", sha[:16])


if __name__ == "__main__":
    main()
