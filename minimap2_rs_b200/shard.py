"""Read sharding across the GPUs of one box: reads are independent (main.rs:193-219 touches only
&Index), so every rank maps a contiguous range of reads against its own replica of the index and the PAF lines are
concatenated in input order.  There is no collective on the data path; torch.distributed is used only to hand the
text back to rank 0."""
import numpy as np


def shard_reads(offs, world, rank):
    """contiguous read range [lo, hi) of `rank`, balanced by bases (offs = read offsets, len nreads+1)"""
    offs = np.asarray(offs, dtype=np.uint64)
    n = offs.size - 1
    if n <= 0:
        return 0, 0
    total = int(offs[-1] - offs[0])
    base = int(offs[0])
    cuts = [0]
    for r in range(1, world):
        target = base + (total * r) // world
        cuts.append(int(np.searchsorted(offs, np.uint64(target), side="left")))
    cuts.append(n)
    cuts = np.maximum.accumulate(np.minimum(cuts, n))
    return int(cuts[rank]), int(cuts[rank + 1])


def map_sharded(map_fn, cat, offs, names, world, rank):
    """run map_fn(cat_slice, offs_slice, names_slice) -> list of PAF lines on this rank's shard"""
    lo, hi = shard_reads(offs, world, rank)
    offs = np.asarray(offs, dtype=np.uint64)
    b0, b1 = int(offs[lo]), int(offs[hi])
    sub_offs = offs[lo:hi + 1] - offs[lo]
    return map_fn(cat[b0:b1], sub_offs, names[lo:hi]), (lo, hi)


def gather_lines(lines, dst=0, group=None):
    """concatenate every rank's PAF lines on rank `dst` in rank (= input) order; other ranks get None"""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return list(lines)
    world = dist.get_world_size(group)
    buf = [None] * world if dist.get_rank(group) == dst else None
    dist.gather_object(list(lines), buf, dst=dst, group=group)
    if buf is None:
        return None
    out = []
    for part in buf:
        out.extend(part)
    return out
