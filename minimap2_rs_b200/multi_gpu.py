"""Bucket-partitioned multi-GPU index build (index.rs:427-475 split over the GPUs of one box).

Phase 1  each rank sketches a contiguous range of sequences (balanced by bases) and sorts its minimizers bucket-major;
Phase 2  all-to-all: every record goes to the owner of its bucket, rank r owning buckets [ceil(r*2^b/R), ceil((r+1)*2^b/R));
Phase 3  each rank groups the buckets it owns (stable re-sort + run-length grouping, index.rs:74-109);
Phase 4  the finished ranges are replicated: per-rank key/value/position arrays are broadcast into place, the per-bucket
         offset tables and the occurrence histogram are summed (all-reduce) — a rank's local offset table is 0 below its
         range and n_local above it, so the element-wise sum over ranks IS the global offset table;
Phase 5  every rank assembles the replicated index (+ its lookup table) and can map reads against it.

The collectives are torch.distributed (NCCL over NVLink); `build_index_sharded_emulated` runs the same library calls for
R virtual ranks inside one process and does the exchange with tensor copies (used by the single-GPU tests).
"""
import numpy as np

from . import Index, shard


def _kroundup64(x):
    x -= 1
    for s in (1, 2, 4, 8, 16, 32):
        x |= x >> s
    return x + 1


def _word_range(offs, lo, hi, total_words):
    """S words touched by sequences [lo, hi) (boundary words are computed by both neighbours, identically)"""
    if hi <= lo:
        return 0, 0
    return int(offs[lo]) // 8, min(total_words, (int(offs[hi]) + 7) // 8)


def build_index_sharded(ctx, cat, offs, names, w=10, k=15, b=14, flag=0, group=None):
    """collective over `group` (default: world); every rank passes the same host arrays and gets the full index"""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device("cuda", ctx.device)
    offs = np.ascontiguousarray(offs, dtype=np.uint64)
    nseq = offs.size - 1
    total = int(offs[-1])
    lo, hi = shard.shard_reads(offs, world, rank)
    # ---- phase 1
    counts = ctx.mg_sketch_sort(cat, offs, lo, hi, w, k, b, flag, world)
    n_local = int(counts.sum())
    send_c = torch.empty(max(1, n_local), dtype=torch.int64, device=dev)
    send_y = torch.empty(max(1, n_local), dtype=torch.int64, device=dev)
    ctx.mg_export_sorted(send_c.data_ptr(), send_y.data_ptr(), n_local)
    # ---- phase 2
    t_counts = torch.from_numpy(counts.astype(np.int64)).to(dev)
    r_counts = torch.empty_like(t_counts)
    dist.all_to_all_single(r_counts, t_counts, group=group)
    in_splits = counts.astype(np.int64).tolist()
    out_splits = r_counts.cpu().tolist()
    n_recv = int(sum(out_splits))
    recv_c = torch.empty(max(1, n_recv), dtype=torch.int64, device=dev)
    recv_y = torch.empty(max(1, n_recv), dtype=torch.int64, device=dev)
    dist.all_to_all_single(recv_c[:n_recv], send_c[:n_local], out_splits, in_splits, group=group)
    dist.all_to_all_single(recv_y[:n_recv], send_y[:n_local], out_splits, in_splits, group=group)
    torch.cuda.synchronize(dev)
    del send_c, send_y
    # ---- phase 3
    part = ctx.mg_build_partial(recv_c.data_ptr(), recv_y.data_ptr(), n_recv, w, k, b, flag)
    del recv_c, recv_y
    raw, hist, big = part.raw()
    # ---- phase 4
    sizes = torch.tensor([raw.n_keys, raw.n_p], dtype=torch.int64, device=dev)
    all_sizes = [torch.empty_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = [t.cpu().tolist() for t in all_sizes]
    nk = [s[0] for s in all_sizes]
    npp = [s[1] for s in all_sizes]
    tk, tp = sum(nk), sum(npp)
    hk = torch.empty(max(1, tk), dtype=torch.int64, device=dev)
    hv = torch.empty(max(1, tk), dtype=torch.int64, device=dev)
    pp = torch.empty(max(1, tp), dtype=torch.int64, device=dev)
    ko, po = int(sum(nk[:rank])), int(sum(npp[:rank]))
    if raw.n_keys:
        ctx.device_copy(hk.data_ptr() + 8 * ko, raw.hkeys, 8 * raw.n_keys)
        ctx.device_copy(hv.data_ptr() + 8 * ko, raw.hvals, 8 * raw.n_keys)
    if raw.n_p:
        ctx.device_copy(pp.data_ptr() + 8 * po, raw.p, 8 * raw.n_p)
    o_k = o_p = 0
    for r in range(world):
        src = dist.get_global_rank(group, r) if group is not None else r
        if nk[r]:
            dist.broadcast(hk[o_k:o_k + nk[r]], src=src, group=group)
            dist.broadcast(hv[o_k:o_k + nk[r]], src=src, group=group)
        if npp[r]:
            dist.broadcast(pp[o_p:o_p + npp[r]], src=src, group=group)
        o_k += nk[r]
        o_p += npp[r]
    nb = 1 << b
    koff = torch.empty(nb + 1, dtype=torch.int64, device=dev)
    poff = torch.empty(nb + 1, dtype=torch.int64, device=dev)
    ctx.device_copy(koff.data_ptr(), raw.bkt_koff, 8 * (nb + 1))
    ctx.device_copy(poff.data_ptr(), raw.bkt_poff, 8 * (nb + 1))
    dist.all_reduce(koff, group=group)
    dist.all_reduce(poff, group=group)
    t_hist = torch.from_numpy(hist.astype(np.int64)).to(dev)
    dist.all_reduce(t_hist, group=group)
    bigs = [None] * world
    dist.all_gather_object(bigs, big.tolist(), group=group)
    big_all = np.array([x for part_ in bigs for x in part_], dtype=np.uint32)
    # 4-bit sequence array: every rank packs the words of its own sequences, then the ranges are broadcast
    words_used = (total + 7) // 8
    s_alloc = _kroundup64(words_used) if total else 0
    S = torch.zeros(max(1, s_alloc), dtype=torch.int32, device=dev)
    torch.cuda.current_stream(dev).synchronize()   # the zero fill runs on torch's stream, mg_pack_seq on the context's
    ranges = [_word_range(offs, *shard.shard_reads(offs, world, r), words_used) for r in range(world)]
    w0, w1 = ranges[rank]
    ctx.mg_pack_seq(cat, total, w0, w1, S.data_ptr())
    for r in range(world):
        a0, a1 = ranges[r]
        if a1 > a0:
            dist.broadcast(S[a0:a1], src=(dist.get_global_rank(group, r) if group is not None else r), group=group)
    torch.cuda.synchronize(dev)
    # ---- phase 5
    idx = ctx.index_assemble(offs, names, w, k, b, flag, tk, tp, hk.data_ptr(), hv.data_ptr(), pp.data_ptr(), koff.data_ptr(),
                             poff.data_ptr(), S.data_ptr(), s_alloc, t_hist.cpu().numpy().astype(np.uint64), big_all)
    part.close()
    return idx


def build_index_sharded_emulated(ctx, cat, offs, names, w=10, k=15, b=14, flag=0, world=2):
    """the same five phases for `world` virtual ranks on ONE GPU / one process (exchange = tensor copies)"""
    import torch
    dev = torch.device("cuda", ctx.device)
    offs = np.ascontiguousarray(offs, dtype=np.uint64)
    total = int(offs[-1])
    send = []
    for rank in range(world):
        lo, hi = shard.shard_reads(offs, world, rank)
        counts = ctx.mg_sketch_sort(cat, offs, lo, hi, w, k, b, flag, world)
        n_local = int(counts.sum())
        c = torch.empty(max(1, n_local), dtype=torch.int64, device=dev)
        y = torch.empty(max(1, n_local), dtype=torch.int64, device=dev)
        ctx.mg_export_sorted(c.data_ptr(), y.data_ptr(), n_local)
        send.append((counts.astype(np.int64), c, y))
    parts = []
    for rank in range(world):  # all-to-all: concatenate, in source-rank order, the slice every source holds for `rank`
        cs, ys = [], []
        for src in range(world):
            counts, c, y = send[src]
            o = int(counts[:rank].sum())
            cs.append(c[o:o + int(counts[rank])])
            ys.append(y[o:o + int(counts[rank])])
        rc, ry = torch.cat(cs), torch.cat(ys)
        n_recv = int(rc.numel())
        rc = rc if n_recv else torch.zeros(1, dtype=torch.int64, device=dev)
        ry = ry if n_recv else torch.zeros(1, dtype=torch.int64, device=dev)
        torch.cuda.synchronize(dev)
        parts.append(ctx.mg_build_partial(rc.data_ptr(), ry.data_ptr(), n_recv, w, k, b, flag))
    raws = [p.raw() for p in parts]
    tk = sum(r[0].n_keys for r in raws)
    tp = sum(r[0].n_p for r in raws)
    hk = torch.empty(max(1, tk), dtype=torch.int64, device=dev)
    hv = torch.empty(max(1, tk), dtype=torch.int64, device=dev)
    pp = torch.empty(max(1, tp), dtype=torch.int64, device=dev)
    nb = 1 << b
    koff = torch.zeros(nb + 1, dtype=torch.int64, device=dev)
    poff = torch.zeros(nb + 1, dtype=torch.int64, device=dev)
    hist = np.zeros(65536, dtype=np.uint64)
    bigs = []
    o_k = o_p = 0
    tmp = torch.empty(nb + 1, dtype=torch.int64, device=dev)
    for raw, h, big in raws:
        if raw.n_keys:
            ctx.device_copy(hk.data_ptr() + 8 * o_k, raw.hkeys, 8 * raw.n_keys)
            ctx.device_copy(hv.data_ptr() + 8 * o_k, raw.hvals, 8 * raw.n_keys)
        if raw.n_p:
            ctx.device_copy(pp.data_ptr() + 8 * o_p, raw.p, 8 * raw.n_p)
        o_k += raw.n_keys
        o_p += raw.n_p
        ctx.device_copy(tmp.data_ptr(), raw.bkt_koff, 8 * (nb + 1))
        koff += tmp
        ctx.device_copy(tmp.data_ptr(), raw.bkt_poff, 8 * (nb + 1))
        poff += tmp
        hist += h
        bigs.extend(big.tolist())
    words_used = (total + 7) // 8
    s_alloc = _kroundup64(words_used) if total else 0
    S = torch.zeros(max(1, s_alloc), dtype=torch.int32, device=dev)
    torch.cuda.current_stream(dev).synchronize()   # see build_index_sharded
    for rank in range(world):
        w0, w1 = _word_range(offs, *shard.shard_reads(offs, world, rank), words_used)
        ctx.mg_pack_seq(cat, total, w0, w1, S.data_ptr())
    torch.cuda.synchronize(dev)
    idx = ctx.index_assemble(offs, names, w, k, b, flag, tk, tp, hk.data_ptr(), hv.data_ptr(), pp.data_ptr(), koff.data_ptr(),
                             poff.data_ptr(), S.data_ptr(), s_alloc, hist, np.array(bigs, dtype=np.uint32))
    for p in parts:
        p.close()
    return idx
