"""The BASELINE.json configs as seeded synthetic workloads, shared by bench.py, tools/config_parity.py and
the tests so that the GPU path and the CPU oracle always see identical bytes.

  C1  configs[0]  145,138,636-bp `chr8`, k=15 w=10, one 600-bp read with 4 substitutions
  C2  configs[1]  same genome, 100k x 10 kb ONT-like reads (3.33 % sub / ins / del each)
  C3  configs[2]  3.1 Gbp: 16 x 193.75 Mbp at even rids, 1-bp N records at odd rids, 0.1 % of positions in N runs
  C4  configs[3]  1 M x 15 kb HiFi-like reads (0.2 / 0.15 / 0.15 %) vs the C3 genome, k=19 w=10
  C5  configs[4]  145 Mbp repeat-rich genome (40 % tandem arrays, 20 % dispersed families), 50k x 100 kb reads at 8 % error
"""
import numpy as np

from tools import gen

C1_LEN = 145_138_636
SEED_C2_GENOME, SEED_C2_READS = 0xB2000002, 0xB2001002
SEED_C3_GENOME, SEED_C4_READS = 0xB2000003, 0xB2001004
SEED_C5_GENOME, SEED_C5_READS = 0xB2000005, 0xB2001005
ERR = {"c2": (0.0333, 0.0333, 0.0333), "c4": (0.002, 0.0015, 0.0015), "c5": (0.027, 0.027, 0.026)}
WK = {"c1": (10, 15), "c2": (10, 15), "c3": (10, 15), "c4": (10, 19), "c5": (10, 15)}


def genome_c2(length=C1_LEN, out=None):
    g = gen.genome(SEED_C2_GENOME, length, out=out)
    return g, np.array([0, length], dtype=np.uint64), ["chr8"]


def read_c1(g):
    """the 600-bp read of configs[0]: genome[5,999,340 .. 5,999,940) with 4 substitutions"""
    q = bytearray(g[5_999_340:5_999_940].tobytes())
    for p in (57, 211, 388, 540):
        q[p] = ord("ACGT"[("ACGT".index(chr(q[p])) + 1) % 4])
    return bytes(q)


def genome_c3(chroms=16, chrom_bp=193_750_000, alloc=None):
    """-> (cat, offs, names).  alloc(nbytes) -> uint8 array (e.g. a pinned buffer); default numpy"""
    names, lens = [], []
    for c in range(chroms):
        names.append("chr%d" % (c + 1))
        lens.append(chrom_bp)
        if c + 1 < chroms:
            names.append("pad%d" % (c + 1))
            lens.append(1)
    offs = np.zeros(len(lens) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(lens)
    total = int(offs[-1])
    cat = alloc(total) if alloc else np.empty(total, dtype=np.uint8)
    r = 0
    for c in range(chroms):
        lo = int(offs[r])
        gen.genome(SEED_C3_GENOME + c, chrom_bp, 1e-3 / 50, 50.0, out=cat[lo:lo + chrom_bp])
        r += 1
        if c + 1 < chroms:
            cat[int(offs[r])] = ord("N")
            r += 1
    return cat, offs, names


def genome_c5(length=C1_LEN):
    g = gen.repeat_genome(SEED_C5_GENOME, length, 0.4, 0.2)
    return g, np.array([0, length], dtype=np.uint64), ["rep"]


def reads(cfg, genome_cat, goffs, nreads, read_len, first=0, out=None):
    """reads [first, first + nreads) of the config's read set (read i depends only on (seed, i), so any rank can generate its
    own shard)"""
    seed = {"c2": SEED_C2_READS, "c4": SEED_C4_READS, "c5": SEED_C5_READS}[cfg]
    return gen.reads(seed, genome_cat, goffs, nreads, read_len, *ERR[cfg], out=out, first=first)


SHAPES = {"c2": dict(reads=100_000, read_len=10_000), "c4": dict(reads=1_000_000, read_len=15_000),
          "c5": dict(reads=50_000, read_len=100_000)}
