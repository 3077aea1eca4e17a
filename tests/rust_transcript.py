"""A SECOND, independent reading of the reference: a literal Python transcription of the Rust sources, statement by
statement (same loops, same variable names), written from /root/reference/src/*.rs and NOT from oracle/mm2_oracle.cpp.

TEST INFRASTRUCTURE ONLY.  tests/test_transcript_cpu.py runs the C++ oracle and this file on the same seeded inputs; a
disagreement means one of the two readings of the Rust text is wrong.  (The reference ships no tests or vectors and
cannot be compiled in this image, SURVEY.md F1/F2, so two independent transcriptions agreeing is the strongest pin
available offline.)

Integer semantics are emulated explicitly: `u64` wrapping through masks, `as i32` through _i32(), `i32 as u64` sign
extension through _i32_as_u64(); f32 arithmetic goes through numpy.float32 one operation at a time (no contraction) and
the two transcendental calls (`f32::ln`, `f32::powf`) through glibc's logf / powf, which is what Rust's std calls.
Pure-Python loops: small inputs only.
"""
import ctypes
import ctypes.util

import numpy as np

_libm = ctypes.CDLL(ctypes.util.find_library("m"))
_libm.logf.restype = ctypes.c_float
_libm.logf.argtypes = [ctypes.c_float]
_libm.powf.restype = ctypes.c_float
_libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]

F32 = np.float32
U64 = (1 << 64) - 1
U64_MAX = U64
I32_MAX = 2147483647


def _i32(x):
    """`x as i32` for an integer: keep the low 32 bits, reinterpret as signed"""
    x &= 0xFFFFFFFF
    return x - (1 << 32) if x & 0x80000000 else x


def _i32_as_u64(x):
    """`x as u64` for an i32: sign extension"""
    return x & U64


def _f32_as_i32(x):
    """`x as i32` for an f32: truncate toward zero, saturate, NaN -> 0"""
    x = float(x)
    if x != x:
        return 0
    if x >= 2147483648.0:
        return I32_MAX
    if x <= -2147483648.0:
        return -2147483648
    return int(x)


def _f32_as_usize(x):
    x = float(x)
    if x != x or x <= 0.0:
        return 0
    if x >= 18446744073709551616.0:
        return U64
    return int(x)


# ---- nt4.rs ---------------------------------------------------------------------------------------------------------------
def nt4(b):
    if b in (65, 97):
        return 0
    if b in (67, 99):
        return 1
    if b in (71, 103):
        return 2
    if b in (84, 116):
        return 3
    return 4


# ---- sketch.rs -----------------------------------------------------------------------------------------------------------
def hash64(key, mask):
    key = ((~key & U64) + ((key << 21) & U64)) & U64 & mask
    key ^= key >> 24
    key = ((key + ((key << 3) & U64) + ((key << 8) & U64)) & U64) & mask
    key ^= key >> 14
    key = ((key + ((key << 2) & U64) + ((key << 4) & U64)) & U64) & mask
    key ^= key >> 28
    key = ((key + ((key << 31) & U64)) & U64) & mask
    return key


class TinyQueue:
    def __init__(self):
        self.front = 0
        self.count = 0
        self.a = [0] * 32

    def clear(self):
        self.front = 0
        self.count = 0

    def push(self, x):
        idx = (self.count + self.front) & 0x1f
        self.a[idx] = x
        self.count += 1

    def shift(self):
        if self.count == 0:
            return -1
        x = self.a[self.front]
        self.front = (self.front + 1) & 0x1f
        self.count -= 1
        return x


def sketch_sequence(seq, w, k, rid, is_hpc, out):
    """sketch.rs:29-100; out: list of (key_span, rid_pos_strand)"""
    assert len(seq) > 0
    assert 0 < w < 256
    assert 0 < k <= 28
    shift1 = 2 * (k - 1)
    mask = (1 << (2 * k)) - 1
    kmer = [0, 0]
    l = 0
    buf_pos = 0
    min_pos = 0
    kmer_span = 0
    buf = [(U64_MAX, U64_MAX)] * w
    mn = (U64_MAX, U64_MAX)
    tq = TinyQueue()
    n = len(seq)
    for i in range(n):
        c = nt4(seq[i])
        info = (U64_MAX, U64_MAX)
        if c < 4:
            if is_hpc:
                skip_len = 1
                if i + 1 < n and nt4(seq[i + 1]) == c:
                    t = i + 2
                    while t < n and nt4(seq[t]) == c:
                        t += 1
                    skip_len = t - i
                tq.push(skip_len)
                kmer_span += skip_len
                if tq.count > k:
                    kmer_span -= tq.shift()
            else:
                kmer_span = l + 1 if l + 1 < k else k
            kmer[0] = ((kmer[0] << 2) | c) & mask
            kmer[1] = (kmer[1] >> 2) | ((3 ^ c) << shift1)
            if kmer[0] != kmer[1]:
                z = 0 if kmer[0] < kmer[1] else 1
                l += 1
                if l >= k and kmer_span < 256:
                    key_span = ((hash64(kmer[z], mask) << 8) & U64) | kmer_span
                    rid_pos_strand = ((rid << 32) | (i << 1) | z) & U64
                    info = (key_span, rid_pos_strand)
        else:
            l = 0
            tq.clear()
            kmer_span = 0
        buf[buf_pos] = info
        if l == w + k - 1 and mn[0] != U64_MAX:
            for j in range(buf_pos + 1, w):
                if mn[0] == buf[j][0] and buf[j][1] != mn[1]:
                    out.append(buf[j])
            for j in range(0, buf_pos):
                if mn[0] == buf[j][0] and buf[j][1] != mn[1]:
                    out.append(buf[j])
        if info[0] <= mn[0]:
            if l >= w + k and mn[0] != U64_MAX:
                out.append(mn)
            mn = info
            min_pos = buf_pos
        elif buf_pos == min_pos:
            if l >= w + k - 1 and mn[0] != U64_MAX:
                out.append(mn)
            mn = (U64_MAX, mn[1])
            for j in range(buf_pos + 1, w):
                if mn[0] >= buf[j][0]:
                    mn = buf[j]
                    min_pos = j
            for j in range(0, buf_pos + 1):
                if mn[0] >= buf[j][0]:
                    mn = buf[j]
                    min_pos = j
            if l >= w + k - 1 and mn[0] != U64_MAX:
                for j in range(buf_pos + 1, w):
                    if mn[0] == buf[j][0] and mn[1] != buf[j][1]:
                        out.append(buf[j])
                for j in range(0, buf_pos + 1):
                    if mn[0] == buf[j][0] and mn[1] != buf[j][1]:
                        out.append(buf[j])
        buf_pos += 1
        if buf_pos == w:
            buf_pos = 0
    if mn[0] != U64_MAX:
        out.append(mn)


# ---- index.rs ------------------------------------------------------------------------------------------------------------
def kroundup64(x):
    x -= 1
    x |= x >> 1
    x |= x >> 2
    x |= x >> 4
    x |= x >> 8
    x |= x >> 16
    x |= x >> 32
    return x + 1


class Bucket:
    def __init__(self):
        self.a = []
        self.p = []
        self.h = None


class Index:
    """index.rs:33-154 + build_index_from_fasta (:427-475) from in-memory records [(name or None, bytes)]"""

    def __init__(self, w, k, b, flag):
        self.w, self.k, self.b, self.flag = w, k, b, flag
        self.n_seq = 0
        self.seq = []   # (name, offset, len, is_alt)
        self.S = []
        self.B = [Bucket() for _ in range(1 << b)]

    @classmethod
    def build(cls, records, w, k, b, flag):
        idx = cls(w, k, b, flag)
        idx.n_seq = len(records)
        is_hpc = (flag & 1) != 0
        minis_by_seq = []
        for rid, (_name, seq) in enumerate(records):
            a = []
            if len(seq) > 0:
                sketch_sequence(seq, w, k, rid, is_hpc, a)
            minis_by_seq.append(a)
        total_len = sum(len(s) for _, s in records)
        words = kroundup64((total_len + 7) // 8) if total_len else 0   # kroundup64(0) wraps to 0 in release builds
        idx.S = [0] * words
        sum_len = 0
        for rid, (name, seq) in enumerate(records):
            for j, ch in enumerate(seq):
                c = nt4(ch)
                o = sum_len + j
                i = o >> 3
                shift = (o & 7) << 2
                v = idx.S[i]
                idx.S[i] = (v & ~(0xF << shift) & 0xFFFFFFFF) | ((c & 0xF) << shift)
            idx.seq.append((name, sum_len, len(seq), False))
            idx.add_minimizers(minis_by_seq[rid])
            sum_len += len(seq)
        idx.post_process()
        return idx

    def add_minimizers(self, v):
        mask = (1 << self.b) - 1
        for m in v:
            self.B[(m[0] >> 8) & mask].a.append(m)

    def post_process(self):
        b_bits = self.b
        for b in self.B:
            if not b.a:
                continue
            b.a.sort(key=lambda x: x[0] >> 8)   # sort_by_key: stable
            n = 1
            total_p = 0
            for j in range(1, len(b.a) + 1):
                if j == len(b.a) or (b.a[j][0] >> 8) != (b.a[j - 1][0] >> 8):
                    if n > 1:
                        total_p += n
                    n = 1
                else:
                    n += 1
            b.p = [0] * total_p
            h = {}
            n = 1
            start_a = 0
            start_p = 0
            for j in range(1, len(b.a) + 1):
                if j == len(b.a) or (b.a[j][0] >> 8) != (b.a[j - 1][0] >> 8):
                    p = b.a[j - 1]
                    key_top = ((p[0] >> 8) >> b_bits) << 1
                    if n == 1:
                        h[key_top | 1] = p[1]
                    else:
                        for kk in range(n):
                            b.p[start_p + kk] = b.a[start_a + kk][1]
                        b.p[start_p:start_p + n] = sorted(b.p[start_p:start_p + n])
                        h[key_top] = (start_p << 32) | n
                        start_p += n
                    start_a = j
                    n = 1
                else:
                    n += 1
            b.h = h
            b.a = []

    def stats(self):
        n_keys = 0
        sum_occ = 0
        for b in self.B:
            if b.h is not None:
                for k_, v in b.h.items():
                    n_keys += 1
                    sum_occ += 1 if (k_ & 1) == 1 else (v & 0xffffffff)
        total_len = sum(s[2] for s in self.seq)
        avg_occ = sum_occ / n_keys if n_keys > 0 else 0.0
        avg_spacing = total_len / sum_occ if sum_occ > 0 else 0.0
        return n_keys, avg_occ, avg_spacing, total_len

    def calc_mid_occ(self, frac):
        counts = []
        for b in self.B:
            if b.h is not None:
                for k_, v in b.h.items():
                    counts.append(1 if (k_ & 1) == 1 else (v & 0xffffffff))
        if not counts:
            return I32_MAX
        counts.sort()
        n = len(counts)
        x = (1.0 - float(F32(frac))) * float(n)   # frac as f64: the f32 value widened
        idx = 0 if x <= 0.0 else int(x)
        idx = min(idx, n - 1)
        return _i32(counts[idx]) + 1

    def get(self, minier):
        """-> None | ('S', y) | ('M', [y...])"""
        mask = (1 << self.b) - 1
        b = self.B[minier & mask]
        if b.h is None:
            return None
        key = (minier >> self.b) << 1
        if (key | 1) in b.h:
            return ("S", b.h[key | 1])
        if key in b.h:
            val = b.h[key]
            off = val >> 32
            n = val & 0xffffffff
            return ("M", b.p[off:off + n])
        return None


# ---- seeds.rs ------------------------------------------------------------------------------------------------------------
def collect_query_minimizers(seq, w, k):
    v = []
    sketch_sequence(seq, w, k, 0, False, v)
    return v


def filter_query_minimizers(mv, q_occ_max, q_occ_frac):
    """in place on the list `mv`"""
    if len(mv) == 0 or float(F32(q_occ_frac)) <= 0.0 or q_occ_max <= 0:
        return
    if _i32(len(mv)) <= q_occ_max:
        return
    keys = [((m[0] >> 8), i) for i, m in enumerate(mv)]
    keys.sort(key=lambda x: x[0])   # sort_unstable_by_key: only the grouping of equal keys is used below
    keep = [True] * len(mv)
    st = 0
    n = len(keys)
    cutoff = _f32_as_usize(F32(len(mv)) * F32(q_occ_frac))
    for i in range(1, n + 1):
        if i == n or keys[i][0] != keys[st][0]:
            cnt = i - st
            if _i32(cnt) > q_occ_max and cnt > cutoff:
                for j in range(st, i):
                    keep[keys[j][1]] = False
            st = i
    j = 0
    for i in range(len(mv)):
        if keep[i]:
            mv[j] = mv[i]
            j += 1
    del mv[j:]


def push_anchor(out, r, m, qlen):
    rid = (r >> 32) & 0xffffffff
    rpos = _i32((r >> 1) & 0xffffffff)
    rstrand = r & 1
    qpos = _i32((m[1] >> 1) & 0xffffffff)
    qstrand = m[1] & 1
    qspan = m[0] & 0xff
    forward = rstrand == qstrand
    if forward:
        x = ((rid << 32) & U64) | _i32_as_u64(rpos)
    else:
        x = (1 << 63) | ((rid << 32) & U64) | _i32_as_u64(rpos)
    if forward:
        y = (qspan << 32) | _i32_as_u64(qpos)
    else:
        qp = _i32_as_u64(_i32(qlen - (qpos + 1 - qspan) - 1))
        y = (qspan << 32) | qp
    out.append((x, y))


def build_anchors_filtered(idx, mv, qlen, mid_occ):
    a = []
    for m in mv:
        minier = m[0] >> 8
        occ = idx.get(minier)
        if occ is not None:
            if occ[0] == "S":
                push_anchor(a, occ[1], m, qlen)
            else:
                if _i32(len(occ[1])) > mid_occ:
                    continue
                for r in occ[1]:
                    push_anchor(a, r, m, qlen)
    a.sort()   # by x, then y (stable; equal anchors are identical)
    return a


# ---- lchain.rs -----------------------------------------------------------------------------------------------------------
def qpos(a):
    return _i32(a[1] & 0xffffffff)


def qspan(a):
    return (a[1] >> 32) & 0xff


def rpos(a):
    return _i32(a[0] & 0xffffffff)


def rev(a):
    return (a[0] >> 63) != 0


def rid(a):
    return (a[0] >> 32) & 0x7fffffff


LN_2 = F32(0.6931471805599453)


def mg_log2(x):
    if x <= 1:
        return F32(0.0)
    return F32(_libm.logf(F32(x))) / LN_2


def comput_sc(ai, aj, max_dist_x, max_dist_y, bw, chn_pen_gap, chn_pen_skip):
    dq = qpos(ai) - qpos(aj)
    if dq <= 0 or dq > max_dist_x:
        return None
    dr = rpos(ai) - rpos(aj)
    if dr == 0 or dq > max_dist_y:
        return None
    dd = abs(dr - dq)
    if dd > bw:
        return None
    dg = min(dr, dq)
    q_span = qspan(aj)
    sc = min(q_span, dg)
    if dd != 0 or dg > q_span:
        lin_pen = F32(chn_pen_gap) * F32(dd) + F32(chn_pen_skip) * F32(dg)
        log_pen = mg_log2(dd + 1) if dd >= 1 else F32(0.0)
        sc -= _f32_as_i32(lin_pen + F32(0.5) * log_pen)
    return sc


class ChainParams:
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def clone(self):
        return ChainParams(**self.__dict__)


def default_chain_params(k):
    """main.rs:105-123"""
    chain_gap_scale = F32(0.8)
    chn_pen_gap = F32(0.01) * chain_gap_scale * F32(k)
    return ChainParams(max_dist_x=5000, max_dist_y=5000, bw=500, max_chain_iter=5000, min_chain_score=40, min_cnt=3,
                       chn_pen_gap=chn_pen_gap, chn_pen_skip=F32(0.0), max_chain_skip=25, max_drop=500, bw_long=20000,
                       rmq_rescue_size=1000, rmq_rescue_ratio=F32(0.1))


def chain_dp_all(anchors, p, trace=None):
    """lchain.rs:59-176 -> (chains, scores); trace (optional dict) receives f, v, pprev and the inner-loop cell count"""
    n = len(anchors)
    if n == 0:
        return [], []
    max_dist_x = p.max_dist_x
    max_dist_y = p.max_dist_y
    if max_dist_x < p.bw:
        max_dist_x = p.bw
    if max_dist_y < p.bw:
        max_dist_y = p.bw
    f = [0] * n
    v = [0] * n
    t = [0] * n
    pprev = [-1] * n
    st = 0
    cells = 0
    for i in range(n):
        while st < i and (rid(anchors[st]) != rid(anchors[i]) or rev(anchors[st]) != rev(anchors[i])
                          or rpos(anchors[i]) > rpos(anchors[st]) + max_dist_x):
            st += 1
        max_j = -1
        max_f = qspan(anchors[i])
        start_j = (i - p.max_chain_iter) if i - p.max_chain_iter > st else st
        n_skip = 0
        for j in range(i - 1, start_j - 1, -1):
            cells += 1
            if rid(anchors[j]) != rid(anchors[i]) or rev(anchors[j]) != rev(anchors[i]):
                continue
            sc0 = comput_sc(anchors[i], anchors[j], max_dist_x, max_dist_y, p.bw, p.chn_pen_gap, p.chn_pen_skip)
            if sc0 is not None:
                sc = sc0 + f[j]
                if sc > max_f:
                    max_f = sc
                    max_j = j
                    if n_skip > 0:
                        n_skip -= 1
                elif t[j] == i:
                    n_skip += 1
                    if n_skip > p.max_chain_skip:
                        break
                if pprev[j] >= 0:
                    t[pprev[j]] = i
        f[i] = max_f
        pprev[i] = max_j
        v[i] = v[max_j] if max_j >= 0 and v[max_j] > max_f else max_f
    if trace is not None:
        trace.update(f=list(f), v=list(v), pprev=list(pprev), cells=cells)
    z = [(f[i], i) for i in range(n) if f[i] > 0]
    if not z:
        return [], []
    z.sort(key=lambda x: x[0])   # sort_unstable_by_key: ties by ascending index assumed
    # (the first backtrack pass, lchain.rs:99-126, only sizes Vec::with_capacity and is not transcribed)
    chains = []
    scores = []
    t = [0] * n
    for k in range(len(z) - 1, -1, -1):
        i0 = z[k][1]
        if t[i0] != 0:
            continue
        i = i0
        end_i = -1
        max_s = 0
        max_i = i
        if i >= 0 and t[i] == 0:
            while True:
                t[i] = 2
                end_i = pprev[i]
                s = z[k][0] if end_i < 0 else z[k][0] - f[end_i]
                if s > max_s:
                    max_s = s
                    max_i = end_i
                elif max_s - s > p.max_drop:
                    break
                if not (i >= 0 and t[i] == 0 and end_i >= 0):
                    break
                i = end_i
            ii = i0
            while ii >= 0 and ii != end_i:
                t[ii] = 0
                ii = pprev[ii]
        v_idxs = []
        i = i0
        end_i = max_i
        while i >= 0 and i != end_i:
            v_idxs.append(i)
            t[i] = 1
            i = pprev[i]
        sc = z[k][0] if i < 0 else z[k][0] - f[i]
        if sc >= p.min_chain_score and len(v_idxs) >= p.min_cnt:
            v_idxs.reverse()
            scores.append(sc)
            chains.append(v_idxs)
    if not chains:
        best_i = 0
        for i in range(n):          # max_by_key returns the LAST maximum
            if f[i] >= f[best_i]:
                best_i = i
        v_idxs = []
        i = best_i
        while i >= 0:
            v_idxs.append(i)
            i = pprev[i]
        v_idxs.reverse()
        if v_idxs:
            chains.append(v_idxs)
            scores.append(v[best_i])
    return sort_chains_stable(anchors, chains, scores)


def chain_dp(anchors, p):
    chains, _ = chain_dp_all(anchors, p)
    return chains[0] if chains else []


def chain_qrange(anchors, chain):
    qs = I32_MAX
    qe = -1
    for i in chain:
        a = anchors[i]
        s = qpos(a) - (qspan(a) - 1)
        e = qpos(a) + 1
        if s < qs:
            qs = s
        if e > qe:
            qe = e
    return max(qs, 0), qe


def chain_trange(anchors, chain):
    ts = I32_MAX
    te = -1
    for i in chain:
        a = anchors[i]
        s = rpos(a) - (qspan(a) - 1)
        e = rpos(a) + 1
        if s < ts:
            ts = s
        if e > te:
            te = e
    return max(ts, 0), te


def sort_chains_stable(anchors, chains, scores):
    idxs = list(range(len(chains)))
    idxs.sort(key=lambda i: (-scores[i], chain_qrange(anchors, chains[i])[0], chain_trange(anchors, chains[i])[0]))
    return [list(chains[i]) for i in idxs], [scores[i] for i in idxs]


def select_primary_secondary(anchors, chains, scores, mask_level):
    primaries = []
    is_primary = [True] * len(chains)
    for ci, chain in enumerate(chains):
        qs, qe = chain_qrange(anchors, chain)
        overlapped = False
        for (_s, (pqs, pqe)) in primaries:
            ov = F32(max(min(qe, pqe) - max(qs, pqs), 0))
            ln = F32(max(qe - qs, 1))
            if ov / ln >= F32(mask_level):
                overlapped = True
                break
        if overlapped:
            is_primary[ci] = False
        else:
            primaries.append((scores[ci], (qs, qe)))
    return is_primary


def select_and_filter_chains(anchors, chains, scores, mask_level, pri_ratio, best_n):
    if not chains:
        return [], [], [], 0, 0
    chains, scores = sort_chains_stable(anchors, chains, scores)
    is_primary = select_primary_secondary(anchors, chains, scores, mask_level)
    out_chains, out_scores, out_is_primary = [], [], []
    s1 = scores[0]
    s2 = 0
    sec_kept = 0
    for i, chain in enumerate(chains):
        if i == 0:
            out_chains.append(chain)
            out_scores.append(scores[i])
            out_is_primary.append(True)
        else:
            if not is_primary[i]:
                continue
            if F32(scores[i]) >= F32(pri_ratio) * F32(s1):
                if sec_kept < best_n:
                    out_chains.append(chain)
                    out_scores.append(scores[i])
                    out_is_primary.append(False)
                    sec_kept += 1
            if s2 == 0:
                s2 = scores[i]
    return out_chains, out_scores, out_is_primary, s1, s2


def merge_adjacent_chains_with_gap(anchors, chains, max_gap_q, max_gap_t):
    items = [(chain_qrange(anchors, ch)[0], i) for i, ch in enumerate(chains)]
    items.sort(key=lambda x: x[0])   # sort_unstable_by_key: ties by ascending index assumed (F10)
    merged = []
    for (_qs, idx) in items:
        ch = chains[idx]
        if not merged:
            merged.append(list(ch))
            continue
        last = merged[-1]
        a_last = anchors[last[-1]]
        a_first = anchors[ch[0]]
        same = rid(a_last) == rid(a_first) and rev(a_last) == rev(a_first)
        _, last_qe = chain_qrange(anchors, last)
        ch_qs, _ = chain_qrange(anchors, ch)
        _, last_te = chain_trange(anchors, last)
        ch_ts, _ = chain_trange(anchors, ch)
        q_gap = ch_qs - last_qe
        t_gap = ch_ts - last_te
        if same and q_gap >= 0 and t_gap >= 0 and q_gap <= max_gap_q and t_gap <= max_gap_t:
            last.extend(ch)
        else:
            merged.append(list(ch))
    return merged


def chain_query_coverage(anchors, chain):
    qs, qe = chain_qrange(anchors, chain)
    return max(qe - qs, 0)


def rescue_long_join(anchors, chains, scores, p, qlen):
    if not chains:
        return chains, scores
    best_cov = chain_query_coverage(anchors, chains[0])
    uncovered = max(qlen - best_cov, 0)
    rescue = uncovered > p.rmq_rescue_size or F32(best_cov) < F32(qlen) * (F32(1.0) - F32(p.rmq_rescue_ratio))
    if not rescue:
        return chains, scores
    p2 = p.clone()
    p2.bw = p.bw_long
    return chain_dp_all(anchors, p2)


# ---- paf.rs --------------------------------------------------------------------------------------------------------------
def paf_from_chain_with_primary(idx, anchors, chain, qname, qseq, is_primary):
    if not chain:
        return None
    strand = "-" if rev(anchors[chain[0]]) else "+"
    qs = I32_MAX
    qe = -1
    ts = I32_MAX
    te = -1
    cm = 0
    for i in chain:
        a = anchors[i]
        cm += 1
        s = qpos(a) - (qspan(a) - 1)
        e = qpos(a) + 1
        if s < qs:
            qs = s
        if e > qe:
            qe = e
        rs = rpos(a) - (qspan(a) - 1)
        re_ = rpos(a) + 1
        if rs < ts:
            ts = rs
        if re_ > te:
            te = re_
    if qs < 0:
        qs = 0
    if ts < 0:
        ts = 0
    rid0 = (anchors[chain[0]][0] >> 32) & 0x7fffffff
    if rid0 >= len(idx.seq):
        raise IndexError("reference panics: idx.seq[%d] out of bounds (F5)" % rid0)
    tname = idx.seq[rid0][0] if idx.seq[rid0][0] is not None else "*"
    tlen = idx.seq[rid0][2]
    mlen = max(qe - qs, 0)
    blen = max(te - ts, 0)
    mv = collect_query_minimizers(qseq, idx.w, idx.k)
    mini_pos = []
    sum_k = 0
    for m in mv:
        mini_pos.append(_i32((m[1] >> 1) & 0xffffffff))
        sum_k += m[0] & 0xff
    avg_k = F32(sum_k) / F32(len(mv)) if mv else F32(idx.k)
    qlen = len(qseq)

    def qpos_fwd(a):
        qp = qpos(a)
        qs_ = qspan(a)
        return qlen - 1 - (qp + 1 - qs_) if rev(a) else qp

    if strand == "-":
        chain_qs_fwd = [qpos_fwd(anchors[i]) for i in reversed(chain)]
    else:
        chain_qs_fwd = [qpos_fwd(anchors[i]) for i in chain]
    dv = F32(0.0)
    if mini_pos and chain_qs_fwd:
        first = chain_qs_fwd[0]
        # slice::binary_search on a sorted slice: Ok(any index whose element equals `first`), then rewound to the first one
        lo, hi, st = 0, len(mini_pos), None
        while lo < hi:
            mid = lo + (hi - lo) // 2
            if mini_pos[mid] == first:
                st = mid
                break
            if mini_pos[mid] < first:
                lo = mid + 1
            else:
                hi = mid
        if st is not None:
            while st > 0 and mini_pos[st - 1] == first:
                st -= 1
            j = st
            k = 1
            en = st
            n_match = 1
            while j + 1 < len(mini_pos) and k < len(chain_qs_fwd):
                j += 1
                if mini_pos[j] == chain_qs_fwd[k]:
                    n_match += 1
                    en = j
                    k += 1
            n_tot = (en - st) + 1
            r_qs_final = qlen - qe if strand == "-" else qs
            r_qe_final = qlen - qs if strand == "-" else qe
            r_rs, r_re = ts, te
            ak = _f32_as_i32(avg_k)
            if r_qs_final > ak and r_rs > ak:
                n_tot += 1
            if (qlen - r_qe_final) > ak and (tlen - r_re) > ak:
                n_tot += 1
            frac = F32(n_match) / F32(n_tot)
            if frac >= F32(1.0):
                dv = F32(0.0)
            else:
                dv = F32(1.0) - F32(_libm.powf(frac, F32(1.0) / max(avg_k, F32(1.0))))
    return dict(qname=qname, qlen=qlen, qstart=qs & 0xFFFFFFFF, qend=qe & 0xFFFFFFFF, strand=strand, tname=tname, tlen=tlen,
                tstart=ts & 0xFFFFFFFF, tend=te & 0xFFFFFFFF, nm=mlen, blen=blen, mapq=60, tp="P" if is_primary else "S", cm=cm,
                s1=0, s2=0, dv=dv, rl=0)


def write_paf(rec):
    if rec["strand"] == "-":
        qs, qe = (rec["qlen"] - rec["qend"]) & 0xFFFFFFFF, (rec["qlen"] - rec["qstart"]) & 0xFFFFFFFF
    else:
        qs, qe = rec["qstart"], rec["qend"]
    return "%s\t%d\t%d\t%d\t%s\t%s\t%d\t%d\t%d\t%d\t%d\t%d\ttp:A:%s\tcm:i:%d\ts1:i:%d\ts2:i:%d\tdv:f:%.4f\trl:i:%d" % (
        rec["qname"], rec["qlen"], qs, qe, rec["strand"], rec["tname"], rec["tlen"], rec["tstart"], rec["tend"], rec["nm"],
        rec["blen"], rec["mapq"], rec["tp"], rec["cm"], rec["s1"], rec["s2"], float(rec["dv"]), rec["rl"])


def write_paf_many_with_scores(idx, anchors, chains, top_s1, top_s2, qname, qseq):
    out = []
    for ci, chain in enumerate(chains):
        rec = paf_from_chain_with_primary(idx, anchors, chain, qname, qseq, ci == 0)
        if rec is not None:
            rec["s1"] = max(top_s1, 0)
            rec["s2"] = max(top_s2, 0)
            out.append(write_paf(rec))
    return out


# ---- main.rs:187-219: `mm2rs align` for one query ---------------------------------------------------------------------------
def align_one(idx, qname, q, w=10, k=15, frac_top_repetitive=2e-4, max_gap=5000, bw=None, bw_long=None, min_cnt=3,
              min_chain_score=40, mask_level=0.5, pri_ratio=0.8, best_n=5, stages=None):
    mv = collect_query_minimizers(q, w, k)
    if stages is not None:
        stages["minis"] = list(mv)
    filter_query_minimizers(mv, 10, 0.01)
    mid_occ = idx.calc_mid_occ(frac_top_repetitive)
    if mid_occ < 10:
        mid_occ = 10
    anchors = build_anchors_filtered(idx, mv, _i32(len(q)), mid_occ)
    p = default_chain_params(k)
    p.max_dist_x = max_gap
    p.max_dist_y = max_gap
    p.min_cnt = min_cnt
    p.min_chain_score = min_chain_score
    if bw is not None:
        p.bw = bw
    if bw_long is not None:
        p.bw_long = bw_long
    tr = {} if stages is not None else None
    chains_all, scores_all = chain_dp_all(anchors, p, tr)
    if stages is not None:
        stages.update(kept=list(mv), anchors=list(anchors), mid_occ=mid_occ, trace=tr, chains=chains_all, scores=scores_all)
    lines = []
    if not chains_all:
        chain = chain_dp(anchors, p)
        rec = paf_from_chain_with_primary(idx, anchors, chain, qname, q, True)
        if rec is not None:
            lines.append(write_paf(rec))
    else:
        chains_rescued, scores_rescued = rescue_long_join(anchors, chains_all, scores_all, p, _i32(len(q)))
        chains_merged = merge_adjacent_chains_with_gap(anchors, chains_rescued, p.max_dist_y, p.max_dist_y)
        chains, _scores, _is_pri, s1, s2 = select_and_filter_chains(anchors, chains_merged, scores_rescued, mask_level, pri_ratio, best_n)
        lines.extend(write_paf_many_with_scores(idx, anchors, chains, s1, s2, qname, q))
    return lines
