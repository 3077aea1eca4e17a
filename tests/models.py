"""Pure-numpy/Python models of the data-parallel reformulations used by the CUDA kernels.

They exist so that the *algorithmic* restatement (position-parallel sketch emission, warp-tiled chaining DP)
can be checked against the oracle on CPU, where no GPU is available; the kernels transcribe these models.
"""
import numpy as np

U64MAX = np.uint64(0xFFFFFFFFFFFFFFFF)
_NT4 = np.full(256, 4, dtype=np.uint8)
for _i, _c in enumerate("ACGT"):
    _NT4[ord(_c)] = _i
    _NT4[ord(_c.lower())] = _i


def hash64(key, mask):
    key = key.astype(np.uint64)
    m = np.uint64(mask)
    with np.errstate(over="ignore"):
        key = (~key + (key << np.uint64(21))) & m
        key ^= key >> np.uint64(24)
        key = (key + (key << np.uint64(3)) + (key << np.uint64(8))) & m
        key ^= key >> np.uint64(14)
        key = (key + (key << np.uint64(2)) + (key << np.uint64(4))) & m
        key ^= key >> np.uint64(28)
        key = (key + (key << np.uint64(31))) & m
    return key


def sketch_model(seq, w, k, rid=0):
    """position-parallel formulation of sketch.rs:29-100 (non-HPC, odd k: no palindromic k-mers)."""
    assert k % 2 == 1
    c = _NT4[np.frombuffer(bytes(seq), dtype=np.uint8)].astype(np.int64)
    L = c.size
    valid = c < 4
    # l[i]: consecutive valid bases ending at i
    idx = np.arange(L)
    lastN = np.maximum.accumulate(np.where(valid, -1, idx))
    l = idx - lastN
    c2 = (c & 3).astype(np.uint64)
    fwd = np.zeros(L, dtype=np.uint64)
    rev = np.zeros(L, dtype=np.uint64)
    for t in range(k):  # base at position i-t contributes to digit t (fwd) / digit k-1-t (rev, complemented)
        if t >= L:
            break
        sh = np.zeros(L, dtype=np.uint64)
        sh[t:] = c2[:L - t]
        fwd |= sh << np.uint64(2 * t)
        shc = np.zeros(L, dtype=np.uint64)
        shc[t:] = np.uint64(3) - c2[:L - t]
        rev |= shc << np.uint64(2 * (k - 1 - t))
    mask = (1 << (2 * k)) - 1
    z = (fwd >= rev).astype(np.uint64)  # 0 if fwd<rev else 1 (equal impossible for odd k)
    km = np.where(z == 0, fwd, rev)
    ok = l >= k
    key = np.where(ok, (hash64(km, mask) << np.uint64(8)) | np.uint64(k), U64MAX)
    val = (np.uint64(rid) << np.uint64(32)) | (idx.astype(np.uint64) << np.uint64(1)) | z
    # windowed newest-argmin with multiplicity of the minimum
    m_cur = np.zeros(L, dtype=np.int64)
    cnt = np.zeros(L, dtype=np.int64)
    for i in range(L):
        lo = max(0, i - w + 1)
        win = key[lo:i + 1]
        mk = win.min()
        eq = np.nonzero(win == mk)[0]
        m_cur[i] = lo + eq[-1]
        cnt[i] = eq.size
    out = []
    for i in range(L):
        li = l[i] if valid[i] else 0
        prev = m_cur[i - 1] if i > 0 else -1
        kp = key[prev] if prev >= 0 else U64MAX
        if li == w + k - 1 and kp != U64MAX:
            for j in range(max(0, i - w + 1), i):
                if key[j] == kp and j != prev:
                    out.append((key[j], val[j]))
        if key[i] <= kp:
            if li >= w + k and kp != U64MAX:
                out.append((key[prev], val[prev]))
        elif prev == i - w:
            if li >= w + k - 1 and kp != U64MAX:
                out.append((key[prev], val[prev]))
            cur = m_cur[i]
            if li >= w + k - 1 and key[cur] != U64MAX and cnt[i] > 1:
                for j in range(max(0, i - w + 1), i + 1):
                    if key[j] == key[cur] and j != cur:
                        out.append((key[j], val[j]))
    cur = m_cur[L - 1]
    if key[cur] != U64MAX:
        out.append((key[cur], val[cur]))
    res = np.zeros(len(out), dtype=[("key_span", "<u8"), ("rid_pos_strand", "<u8")])
    if out:
        res["key_span"] = [o[0] for o in out]
        res["rid_pos_strand"] = [o[1] for o in out]
    return res


def chain_fwd_model(anchors, p, half_log_lut=None):
    """warp-tiled (32 predecessors per step) forward DP of lchain.rs:59-92; returns f, pprev, v, cells_tiles."""
    import math
    x = anchors["x"].astype(np.uint64)
    y = anchors["y"].astype(np.uint64)
    n = x.size
    rpos = (x & np.uint64(0xffffffff)).astype(np.uint32).view(np.int32).astype(np.int64)
    hi = (x >> np.uint64(32)).astype(np.int64)  # rev | rid
    qpos = (y & np.uint64(0xffffffff)).astype(np.uint32).view(np.int32).astype(np.int64)
    qspan = ((y >> np.uint64(32)) & np.uint64(0xff)).astype(np.int64)
    max_dist_x = max(p.max_dist_x, p.bw)
    max_dist_y = max(p.max_dist_y, p.bw)
    f = np.zeros(n, dtype=np.int64)
    v = np.zeros(n, dtype=np.int64)
    t = np.zeros(n, dtype=np.int64)
    pprev = np.full(n, -1, dtype=np.int64)
    f32 = np.float32

    def pen(dd, dg):
        lin = f32(f32(p.chn_pen_gap) * f32(dd)) + f32(f32(p.chn_pen_skip) * f32(dg))
        lg = f32(0.0) if dd < 1 else f32(f32(math.log(f32(dd + 1))) if False else np.log(f32(dd + 1), dtype=np.float32)) / f32(0.6931472)
        return int(f32(lin) + f32(0.5) * lg)

    st = 0
    for i in range(n):
        while st < i and (hi[st] != hi[i] or rpos[i] > rpos[st] + max_dist_x):
            st += 1
        max_f = int(qspan[i])
        max_j = -1
        n_skip = 0
        start_j = max(st, i - p.max_chain_iter)
        jb = i - 1
        done = False
        while jb >= start_j and not done:
            lanes = [jb - ln for ln in range(32) if jb - ln >= start_j]
            sc = [None] * len(lanes)
            for q, j in enumerate(lanes):
                if hi[j] != hi[i]:
                    continue
                dq = qpos[i] - qpos[j]
                if dq <= 0 or dq > max_dist_x:
                    continue
                dr = rpos[i] - rpos[j]
                if dr == 0 or dq > max_dist_y:
                    continue
                dd = abs(dr - dq)
                if dd > p.bw:
                    continue
                dg = min(dr, dq)
                s = min(int(qspan[j]), dg)
                if dd != 0 or dg > qspan[j]:
                    s -= pen(int(dd), int(dg))
                sc[q] = int(s + f[j])
            # phase 1: all valid lanes mark (marks past the break point are harmless)
            for q, j in enumerate(lanes):
                if sc[q] is not None and pprev[j] >= 0:
                    t[pprev[j]] = i
            # phase 2: records via exclusive prefix max; n_skip via composition of x -> max(x+a, b)
            run_max = max_f
            rec = [False] * len(lanes)
            for q in range(len(lanes)):
                if sc[q] is not None and sc[q] > run_max:
                    rec[q] = True
                    run_max = sc[q]
            # maps (a, b): x -> max(x + a, b)
            xs = n_skip
            brk = -1
            for q, j in enumerate(lanes):
                if sc[q] is None:
                    continue
                if rec[q]:
                    xs = max(xs - 1, 0)
                elif t[j] == i:
                    xs = xs + 1
                    if xs > p.max_chain_skip:
                        brk = q
                        break
            upto = len(lanes) if brk < 0 else brk
            for q in range(upto):
                if rec[q]:
                    max_f = sc[q]
                    max_j = lanes[q]
            n_skip = xs
            if brk >= 0:
                done = True
            jb -= 32
        f[i] = max_f
        pprev[i] = max_j
        v[i] = v[max_j] if (max_j >= 0 and v[max_j] > max_f) else max_f
    return f, pprev, v


def chain_dense_model(anchors, p, lut=None):
    """CPU model of chain_dense_kernel's tile pipeline (lchain.cu): anchors are processed in tiles of 32.  While the sequential
    warp walks tile T-1, the helper warps evaluate, for every anchor of tile T, the FAR part of its predecessor window
    (j < i0 - 32: final DP state) into per-(anchor, 32-aligned j-tile) summaries {valid bits, best score} and a per-anchor mark
    bitmask (lchain.rs:86).  The sequential pass then visits, in the reference's order, the ring [i-32, i-1], the rest of the
    previous tile [i0-32, i-33], and the far tiles through their summaries (a far tile is re-evaluated cell by cell only when
    its best score beats the running maximum).  Returns f, pprev, v, cells."""
    x = anchors["x"].astype(np.uint64)
    y = anchors["y"].astype(np.uint64)
    n = x.size
    rpos = (x & np.uint64(0xffffffff)).astype(np.uint32).view(np.int32).astype(np.int64)
    hi = (x >> np.uint64(32)).astype(np.int64)
    qpos = (y & np.uint64(0xffffffff)).astype(np.uint32).view(np.int32).astype(np.int64)
    qspan = ((y >> np.uint64(32)) & np.uint64(0xff)).astype(np.int64)
    mdx = max(p.max_dist_x, p.bw)
    mdy = max(p.max_dist_y, p.bw)
    f32 = np.float32
    f = np.zeros(n, dtype=np.int64)
    v = np.zeros(n, dtype=np.int64)
    pprev = np.full(n, -1, dtype=np.int64)
    pen_cache = {}

    def pen(dd, dg):
        key = (dd, dg if p.chn_pen_skip != 0 else 0)
        if key not in pen_cache:
            lin = f32(f32(p.chn_pen_gap) * f32(dd)) + f32(f32(p.chn_pen_skip) * f32(dg))
            lg = f32(0.0) if dd < 1 else np.log(f32(dd + 1), dtype=np.float32) / f32(0.6931472)
            pen_cache[key] = int(f32(lin) + f32(0.5) * lg)
        return pen_cache[key]

    def score(i, j):
        dq = qpos[i] - qpos[j]
        if dq <= 0 or dq > mdx:
            return None
        dr = rpos[i] - rpos[j]
        if dr == 0 or dq > mdy:
            return None
        dd = abs(dr - dq)
        if dd > p.bw:
            return None
        dg = min(dr, dq)
        s = min(int(qspan[j]), dg)
        if dd != 0 or dg > qspan[j]:
            s -= pen(int(dd), int(dg))
        return int(s + f[j])

    # window starts (static)
    st_arr = np.zeros(n, dtype=np.int64)
    st = 0
    for i in range(n):
        while st < i and (hi[st] != hi[i] or rpos[i] > rpos[st] + mdx):
            st += 1
        st_arr[i] = max(st, i - p.max_chain_iter)
    cells = 0
    for i0 in range(0, n, 32):
        far_hi = i0 - 32
        # ---- phase A: summaries of the far part, from DP state that is final when tile i0-32 is still being walked
        summ = {}
        for a in range(min(32, n - i0)):
            i = i0 + a
            lo = int(st_arr[i])
            if far_hi <= 0 or lo >= far_hi:
                continue
            V, TM, MK = {}, {}, set()
            for j in range(lo, far_hi):
                s = score(i, j)
                jt = j >> 5
                if s is not None:
                    V[jt] = V.get(jt, 0) | (1 << (j & 31))
                    TM[jt] = max(TM.get(jt, -(1 << 40)), s)
                    if pprev[j] >= lo:
                        MK.add(int(pprev[j]))
            summ[a] = (lo, V, TM, MK)
        # ---- phase B: the sequential walk
        for a in range(min(32, n - i0)):
            i = i0 + a
            lo = int(st_arr[i])
            max_f = int(qspan[i])
            max_j = -1
            n_skip = 0
            marks = set()
            broke = False

            def visit(j):
                nonlocal max_f, max_j, n_skip, broke, cells
                cells += 1
                s = score(i, j)
                if s is None:
                    return
                if s > max_f:
                    max_f, max_j = s, j
                    if n_skip > 0:
                        n_skip -= 1
                elif j in marks:
                    n_skip += 1
                    if n_skip > p.max_chain_skip:
                        broke = True
                        return
                if pprev[j] >= 0:
                    marks.add(int(pprev[j]))

            # ring + rest of the previous tile: cell by cell
            near_lo = max(lo, far_hi, 0)
            j = i - 1
            while j >= near_lo and not broke:
                visit(j)
                j -= 1
            if not broke and a in summ:
                flo, V, TM, MK = summ[a]
                marks |= MK                      # far marks: every far valid j' marks its predecessor, whatever the break does later
                jt = (far_hi - 1) >> 5
                while jt >= (flo >> 5) and not broke:
                    jlo, jhi = max(flo, jt << 5), min(far_hi, (jt + 1) << 5)
                    if TM.get(jt, -(1 << 40)) > max_f:
                        for j in range(jhi - 1, jlo - 1, -1):
                            visit(j)
                            if broke:
                                break
                    else:
                        # no record in this tile: n_skip only grows, by one per marked valid cell, in visiting order
                        vb = V.get(jt, 0)
                        mk = [j for j in range(jhi - 1, jlo - 1, -1) if (vb >> (j & 31)) & 1 and j in marks]
                        need = max(p.max_chain_skip + 1 - n_skip, 1)
                        if len(mk) >= need:
                            cells += jhi - mk[need - 1]
                            broke = True
                        else:
                            n_skip += len(mk)
                            cells += jhi - jlo
                    jt -= 1
            f[i] = max_f
            pprev[i] = max_j
            v[i] = v[max_j] if (max_j >= 0 and v[max_j] > max_f) else max_f
    return f, pprev, v, cells


def sketch_model_v4(seq, w, k, rid=0, region=2048, ch=8, is_hpc=False):
    """Model of sketch_tile_kernel_v4 (csrc/sketch.cu): every tile owns T = region - 2w POSITIONS of one sequence.

    * clean chunk (no N / sequence START in bases [c-w-k+1, c+ch-1+w]): position x is a minimizer iff its key equals the
      minimum of some full window that holds it, i.e. the morphological opening (window minimum, then window maximum
      of the minima over the next w windows) returns the key itself;
    * dirty chunk: its positions are marked by a literal simulation of the steps of this chunk and the next
      ceil((w + ch - 1) / ch) chunks (the by-step rules of sketch_model), each step seeded from the keys of its window.
    The reference emits positions in ascending order and never twice, so the marked set IS the output."""
    assert k % 2 == 1 and w >= ch + 1
    c = _NT4[np.frombuffer(bytes(seq), dtype=np.uint8)].astype(np.int64)
    L = c.size
    valid = c < 4
    idx = np.arange(L)
    lastN = np.maximum.accumulate(np.where(valid, -1, idx)) if L else idx
    l = idx - lastN
    c2 = (c & 3).astype(np.uint64)
    fwd = np.zeros(L, dtype=np.uint64)
    rev = np.zeros(L, dtype=np.uint64)
    for t in range(min(k, L)):
        sh = np.zeros(L, dtype=np.uint64)
        sh[t:] = c2[:L - t]
        fwd |= sh << np.uint64(2 * t)
        shc = np.zeros(L, dtype=np.uint64)
        shc[t:] = np.uint64(3) - c2[:L - t]
        rev |= shc << np.uint64(2 * (k - 1 - t))
    mask = (1 << (2 * k)) - 1
    z = (fwd >= rev).astype(np.uint64)
    km = np.where(z == 0, fwd, rev)
    # key_span as the reference compares it (sketch.rs:74): hash << 8 | kmer_span.  -H (sketch.rs:51-61, where the loop index
    # is NOT advanced over a run): the k-mers stay those of the plain sequence, kmer_span is the sum over the last k bases of
    # the homopolymer run that remains from each base on, and a k-mer whose span reaches 256 is dropped WITHOUT resetting l
    span = np.full(L, k, dtype=np.int64)
    if is_hpc and L:
        skip = np.ones(L, dtype=np.int64)
        for i in range(L - 2, -1, -1):
            if valid[i] and valid[i + 1] and c[i] == c[i + 1]:
                skip[i] = skip[i + 1] + 1
        cs = np.concatenate([[0], np.cumsum(skip)])
        span = cs[1:] - cs[np.maximum(idx + 1 - k, 0)]
    gkey = np.where((l >= k) & (span < 256), (hash64(km, mask) << np.uint64(8)) | span.astype(np.uint64), U64MAX) if L else np.zeros(0, dtype=np.uint64)
    T = region - 2 * w
    D = (w + ch - 1) // ch
    ntiles = max(1, (L + T - 1) // T)
    marked = np.zeros(L, dtype=bool)

    def K(p):  # key at sequence position p (outside: invalid)
        return gkey[p] if 0 <= p < L else U64MAX

    def isN(b):   # only real N: neither end of the sequence needs the by-step rules (see start_tie below)
        return 0 <= b < L and not valid[b]

    # The first w keys of a sequence: sketch.rs:80-86 treats ties among them specially (duplicates of the first partial
    # minimum are emitted at l == w+k-1, and a minimum replaced before l reaches w+k is not).  With an equal pair there, the
    # chunks that see the sequence start fall back to the by-step rules.
    first = gkey[k - 1:k - 1 + w]
    start_tie = first.size < w or np.unique(first).size < first.size or bool((first == U64MAX).any())
    garbage = np.random.default_rng(L * 131 + w).integers(0, 1 << 40, size=region, dtype=np.uint64)

    for tile in range(ntiles):
        s = tile * T
        P0 = s - w
        nown = min(T, L - s)
        nch = region // ch
        # past the end of the sequence the kernel computes keys from whatever follows in the buffer: nothing may depend on them
        key = np.array([K(P0 + u) if P0 + u < L else garbage[u] for u in range(region)], dtype=np.uint64)   # no k-mer before position k-1: KMAX
        u_first = (k - 1) + (w - 1) - P0              # region index of the first full window
        u_last = min(region - 1, L - 1 - P0)
        has_end = (L - 1 - P0) <= region - 1
        dirty = np.zeros(nch, dtype=bool)
        for t in range(nch):
            c0 = t * ch
            lo, hi = P0 + c0 - w - k + 1, P0 + min(c0 + ch - 1 + w, region - 1)
            dirty[t] = any(isN(b) for b in range(lo, hi + 1)) or (start_tie and lo < 0)
        emit = np.zeros(region, dtype=bool)
        # by-position, clean chunks
        # windows that would run past the last k-mer of the sequence do not exist: M = 0 (below every key) for them
        M = np.array([key[max(0, u - w + 1):u + 1].min() if u_first <= u <= u_last else 0 for u in range(region)], dtype=np.uint64)
        for t in range(nch):
            if dirty[t]:
                continue
            for x in range(t * ch, t * ch + ch):
                if x + w - 1 < region and x - w + 1 >= 0 and P0 + x >= k - 1:
                    if M[x:x + w].max() == key[x] and key[x] != U64MAX:
                        emit[x] = True
        # by-step, chunks within D of a dirty chunk
        for t in range(nch):
            if not dirty[max(0, t - D):t + 1].any():
                continue
            for u in range(t * ch, t * ch + ch):
                if u < w or u > u_last:
                    continue
                i = P0 + u
                li = l[i] if valid[i] else 0
                win = key[u - w:u]                      # window of step u-1: [u-w, u-1]
                mk = win.min()
                prev = u - w + int(np.nonzero(win == mk)[0][-1])
                kp = key[prev]
                ki = key[u]
                win2 = key[u - w + 1:u + 1]
                mk2 = win2.min()
                eq2 = np.nonzero(win2 == mk2)[0]
                cur = u - w + 1 + int(eq2[-1])
                if kp != U64MAX:
                    if li == w + k - 1:
                        for j in range(u - w + 1, u):
                            if key[j] == kp and j != prev:
                                emit[j] = True
                    if ki <= kp:
                        if li >= w + k:
                            emit[prev] = True
                    elif prev == u - w:
                        if li >= w + k - 1:
                            emit[prev] = True
                            if key[cur] != U64MAX and eq2.size > 1:
                                for j in range(u - w + 1, u + 1):
                                    if key[j] == key[cur] and j != cur:
                                        emit[j] = True
                if has_end and u == u_last and key[cur] != U64MAX:
                    emit[cur] = True
        for x in range(w, w + max(nown, 0)):
            if emit[x]:
                marked[P0 + x] = True
    pos = np.nonzero(marked)[0]
    res = np.zeros(pos.size, dtype=[("key_span", "<u8"), ("rid_pos_strand", "<u8")])
    res["key_span"] = gkey[pos]
    res["rid_pos_strand"] = (np.uint64(rid) << np.uint64(32)) | (pos.astype(np.uint64) << np.uint64(1)) | z[pos]
    return res
