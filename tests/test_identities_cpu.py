"""CPU suite: the integer identities the kernels lean on where they deviate from the literal form of the reference's tests
(each kernel states them next to the code; here they are checked on edge values and at random)."""
import numpy as np

I32 = np.int32
EDGE = np.array([-2**31, -2**31 + 1, -5001, -2, -1, 0, 1, 2, 31, 32, 499, 500, 501, 4999, 5000, 5001, 20000, 2**31 - 2, 2**31 - 1], dtype=np.int64)


def _u32(x):
    return (np.asarray(x, dtype=np.int64) & 0xFFFFFFFF).astype(np.uint64)


def _wsub(a, b):
    return ((np.asarray(a, dtype=np.int64) - np.asarray(b, dtype=np.int64) + 2**31) % 2**32 - 2**31).astype(np.int64)


def test_chain_fast_range_compares():
    """lchain.cu chain_sc_flat<.., TRIM>: lchain.rs:18-24 `dq <= 0 || dq > max_dist_x`, `dq > max_dist_y`, `dd > bw` (dd = |dr - dq|
    in wrapping i32, so it can be INT_MIN) as one unsigned compare each, valid for bw >= 0 (mdx, mdy >= bw)"""
    rng = np.random.default_rng(1)
    vals = np.concatenate([EDGE, rng.integers(-2**31, 2**31, 4000), rng.integers(-6000, 6000, 4000)])
    for bw in (0, 1, 500, 20000, 100000, 2**31 - 1):
        for mdx, mdy in ((5000, 5000), (bw, 7), (2**31 - 1, 12)):
            mdx, mdy = max(mdx, bw), max(mdy, bw)
            dq = vals
            lit = (dq > 0) & (dq <= mdx) & (dq <= mdy)
            fast = _u32(_wsub(dq, 1)) < np.uint64(min(mdx, mdy))
            assert (lit == fast).all()
            dd = vals
            assert (((dd <= bw) & (dd >= 0)) == (_u32(dd) <= np.uint64(bw))).all()


def test_ring_slots_pass_max_chain_iter_when_it_is_at_least_32():
    """chain_read<.., FAST>: a filled ring slot holds j in [i - 32, i - 1]; lchain.rs:78 asks j >= i - max_chain_iter"""
    for max_iter in (32, 33, 5000, 2**31 - 1):
        for i in (1, 31, 32, 33, 100000, 2**31 - 1):
            j = np.arange(max(i - 32, 0), i, dtype=np.int64)
            assert (j >= max(int(_wsub(i, max_iter)), 0)).all()


def test_bloom_bit_select_without_indexing():
    """seeds.cu: bit b of a 128-bit block {x, y, z, w} by two selects on bits 5 and 6 of b == word b >> 5, bit b & 31"""
    rng = np.random.default_rng(2)
    blk = rng.integers(0, 2**32, (500, 4), dtype=np.uint64)
    for b in range(128):
        lit = (blk[:, b >> 5] >> np.uint64(b & 31)) & np.uint64(1)
        w01 = blk[:, 1] if b & 32 else blk[:, 0]
        w23 = blk[:, 3] if b & 32 else blk[:, 2]
        sel = ((w23 if b & 64 else w01) >> np.uint64(b & 31)) & np.uint64(1)
        assert (lit == sel).all()


def test_left_aligned_hash64_steps_are_multiplies():
    """sketch.cu sk4_hash: with the key left-aligned in a B-bit register (B = 32 or 64, s = B - 2k free low bits) the `& mask`
    of every sketch.rs:4-13 step is the register's natural wrap, `~key + (key << 21)` is key * (2^21 - 1) - 1 and
    `key ^ key >> n` keeps the low s bits clear when the shifted copy is masked"""
    rng = np.random.default_rng(3)
    for k, B in ((15, 32), (11, 32), (19, 64), (28, 64), (15, 64)):
        bits = 2 * k
        mask = (1 << bits) - 1
        s = B - bits
        wrap = (1 << B) - 1
        himask = wrap ^ ((1 << s) - 1)
        for key in [0, 1, mask, mask - 1] + [int(x) for x in rng.integers(0, mask + 1, 300, dtype=np.uint64)]:
            # the reference (right-aligned, masked)
            h = key
            h = (~h + (h << 21)) & mask
            h = h ^ (h >> 24)
            h = ((h + (h << 3)) + (h << 8)) & mask
            h = h ^ (h >> 14)
            h = ((h + (h << 2)) + (h << 4)) & mask
            h = h ^ (h >> 28)
            h = (h + (h << 31)) & mask
            # left-aligned
            x = (key << s) & wrap
            x = (x * ((1 << 21) - 1) - (1 << s)) & wrap
            x ^= (x >> 24) & himask
            x = (x * 265) & wrap
            x ^= (x >> 14) & himask
            x = (x * 21) & wrap
            x ^= (x >> 28) & himask
            x = (x * ((1 << 31) + 1)) & wrap
            assert x == (h << s) & wrap and x & ((1 << s) - 1) == 0


def test_fine_cdf_is_monotone_and_loop_form_independent():
    """mm2_internal.cuh index_fine_cdf: u = 1 - (1 - x)^(2^pw) in 32-bit fixed point; the lookup relies on it being monotone in
    hk (a fine bucket is a contiguous run of the sorted keys) and on build and lookup computing the same integer"""
    def cdf(hk, R, j, pw, unrolled):
        if j <= 0:
            return 0
        t = ((1 << R) - 1) - hk
        t32 = (t >> (R - 32)) if R >= 32 else ((t << (32 - R)) & 0xFFFFFFFF)
        if unrolled:
            for s in range(5):
                if s < pw:
                    t32 = (t32 * t32) >> 32
        else:
            for _ in range(pw):
                t32 = (t32 * t32) >> 32
        return ((~t32) & 0xFFFFFFFF) >> (32 - j)
    rng = np.random.default_rng(4)
    for R, j, pw in ((16, 11, 2), (24, 15, 2), (24, 9, 0), (42, 30, 5), (6, 3, 1), (16, 1, 3)):
        hk = np.sort(rng.integers(0, 1 << R, 3000, dtype=np.uint64)).tolist() + [(1 << R) - 1]
        hk = [0] + hk
        prev = -1
        for h in hk:
            a, b = cdf(int(h), R, j, pw, True), cdf(int(h), R, j, pw, False)
            assert a == b and a >= prev and a < (1 << j)
            prev = a
