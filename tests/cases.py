"""Seeded inputs shared by the CPU (oracle/model) and GPU parity tests."""
import numpy as np


def rnd_seq(rng, n, alpha="ACGT"):
    return bytes(rng.choice(np.frombuffer(alpha.encode(), dtype=np.uint8), n).astype(np.uint8))


def sketch_cases(seed=1, big=False):
    """list of (name, seq bytes): random, N-spliced, low-complexity, tiny, tile-boundary sized"""
    rng = np.random.default_rng(seed)
    out = []
    out.append(("random3k", rnd_seq(rng, 3000)))
    b = bytearray(rnd_seq(rng, 5000))
    for _ in range(40):
        p = int(rng.integers(0, 4900))
        b[p:p + 1] = b"N"
    out.append(("singleN", bytes(b)))
    for _ in range(15):
        p = int(rng.integers(0, 4900))
        ln = int(rng.integers(1, 60))
        b[p:p + ln] = b"N" * ln
    out.append(("runsN", bytes(b[:5000])))
    out.append(("lowcomplex_AC", rnd_seq(rng, 3000, "AC")))
    out.append(("tandem", b"ACG" * 700 + rnd_seq(rng, 100) + b"A" * 400 + b"AT" * 300 + rnd_seq(rng, 50)))
    out.append(("homopolymer", b"A" * 300))
    out.append(("lowercase_iupac", rnd_seq(rng, 2000, "ACGTacgtNRYn")))
    for n in (1, 5, 14, 15, 16, 24, 25, 26, 40):
        out.append(("tiny%d" % n, rnd_seq(rng, n)))
    # around the tile size of the CUDA kernel (2048 - w steps per tile)
    for n in (2037, 2038, 2039, 2048, 4076, 4077, 6200):
        out.append(("tile%d" % n, rnd_seq(rng, n)))
    s = bytearray(rnd_seq(rng, 9000))
    s[2030:2050] = b"N" * 20          # N run straddling a tile boundary
    s[4070:4080] = s[4060:4070]       # repeat straddling a tile boundary
    out.append(("straddle", bytes(s)))
    if big:
        out.append(("random300k", rnd_seq(rng, 300_000)))
        g = bytearray(rnd_seq(rng, 200_000))
        for _ in range(200):
            p = int(rng.integers(0, 199_000))
            ln = int(rng.integers(1, 120))
            g[p:p + ln] = b"N" * ln
        out.append(("N200k", bytes(g)))
    return out


WK = [(10, 15), (10, 19), (11, 21), (3, 5), (1, 7), (50, 27), (5, 3), (19, 17), (200, 11)]


def cat_offs(seqs):
    offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(s) for s in seqs])
    cat = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    return cat, offs
