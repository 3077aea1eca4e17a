"""Minimal reader / writer of the .mmi byte layout (index.rs:233-307) for the fixtures of the loader tests,
and an independent Python model of klib's khash slot order (what `minimap2 -d` dumps).  Test infrastructure only."""
import struct

import numpy as np


def parse(data):
    assert data[:4] == b"MMI\x02"
    w, k, b, n_seq, flag = struct.unpack_from("<5I", data, 4)
    o = 24
    seqs = []
    total = 0
    for _ in range(n_seq):
        nl = data[o]
        name = data[o + 1:o + 1 + nl]
        o += 1 + nl
        (ln,) = struct.unpack_from("<I", data, o)
        o += 4
        seqs.append((name, ln))
        total += ln
    buckets = []
    for _ in range(1 << b):
        (n,) = struct.unpack_from("<I", data, o)
        o += 4
        p = np.frombuffer(data, dtype="<u8", count=n, offset=o).copy()
        o += 8 * n
        (size,) = struct.unpack_from("<I", data, o)
        o += 4
        ent = np.frombuffer(data, dtype="<u8", count=2 * size, offset=o).reshape(size, 2).copy()
        o += 16 * size
        buckets.append((p, ent))
    S = data[o:]
    return dict(w=w, k=k, b=b, flag=flag, seqs=seqs, buckets=buckets, S=S, total=total)


def serialise(m, with_seq=True):
    out = [b"MMI\x02", struct.pack("<5I", m["w"], m["k"], m["b"], len(m["seqs"]), m["flag"])]
    for name, ln in m["seqs"]:
        out.append(bytes([len(name)]) + name + struct.pack("<I", ln))
    for p, ent in m["buckets"]:
        out.append(struct.pack("<I", len(p)) + np.ascontiguousarray(p, dtype="<u8").tobytes())
        out.append(struct.pack("<I", len(ent)) + np.ascontiguousarray(ent, dtype="<u8").tobytes())
    if with_seq:
        out.append(m["S"])
    return b"".join(out)


class Khash:
    """klib khash.h 0.2.8, KHASH_INIT(idx, uint64_t, uint64_t, 1, idx_hash, idx_eq) with idx_hash(a) = (a) >> 1 (stored in a
    32-bit khint_t) and idx_eq(a, b) = (a >> 1 == b >> 1); quadratic (triangular) probing, load factor 0.77"""
    EMPTY, LIVE, DEL = 0, 1, 2

    def __init__(self):
        self.nb = self.size = self.n_occ = self.upper = 0
        self.keys, self.vals, self.fl = [], [], []

    @staticmethod
    def _roundup32(x):
        x -= 1
        for s in (1, 2, 4, 8, 16):
            x |= x >> s
        return (x + 1) & 0xFFFFFFFF

    def resize(self, want):
        nn = max(4, self._roundup32(want))
        if self.size >= int(nn * 0.77 + 0.5):
            return
        nfl = [self.EMPTY] * nn
        if self.nb < nn:
            self.keys += [0] * (nn - self.nb)
            self.vals += [0] * (nn - self.nb)
        for j in range(self.nb):
            if self.fl[j] != self.LIVE:
                continue
            key, val = self.keys[j], self.vals[j]
            self.fl[j] = self.DEL
            while True:
                i = ((key >> 1) & 0xFFFFFFFF) & (nn - 1)
                step = 0
                while nfl[i] != self.EMPTY:
                    step += 1
                    i = (i + step) & (nn - 1)
                nfl[i] = self.LIVE
                if i < self.nb and self.fl[i] == self.LIVE:
                    self.keys[i], key = key, self.keys[i]
                    self.vals[i], val = val, self.vals[i]
                    self.fl[i] = self.DEL
                else:
                    self.keys[i], self.vals[i] = key, val
                    break
        self.fl = nfl
        self.nb = nn
        self.n_occ = self.size
        self.upper = int(nn * 0.77 + 0.5)

    def put(self, key, val):
        if self.n_occ >= self.upper:
            self.resize(self.nb - 1 if self.nb > (self.size << 1) else self.nb + 1)
        mask = self.nb - 1
        i = ((key >> 1) & 0xFFFFFFFF) & mask
        step = 0
        while self.fl[i] != self.EMPTY:      # no deletions ever happen here, and keys are distinct
            step += 1
            i = (i + step) & mask
        self.keys[i], self.vals[i], self.fl[i] = key, val, self.LIVE
        self.size += 1
        self.n_occ += 1

    def dump(self):
        return [(self.keys[i], self.vals[i]) for i in range(self.nb) if self.fl[i] == self.LIVE]


def khash_order(ent_sorted):
    """entries in ascending key order -> the order mm_idx_dump writes them"""
    h = Khash()
    h.resize(len(ent_sorted))
    for kx, vx in ent_sorted:
        h.put(int(kx), int(vx))
    return h.dump()
