"""CPU suite: the N>1 host logic (read sharding + ordered gather) with gloo, world_size 2."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_reads_partitions_exactly():
    sys.path.insert(0, ROOT)
    from minimap2_rs_b200 import shard
    rng = np.random.default_rng(0)
    for n in (0, 1, 2, 7, 1000):
        lens = rng.integers(1, 20000, n)
        offs = np.zeros(n + 1, dtype=np.uint64)
        offs[1:] = np.cumsum(lens)
        for world in (1, 2, 3, 8):
            cover = []
            for r in range(world):
                lo, hi = shard.shard_reads(offs, world, r)
                assert 0 <= lo <= hi <= n
                cover.extend(range(lo, hi))
                if n >= 1000:
                    assert abs(int(offs[hi] - offs[lo]) - int(offs[-1]) / world) < 25000   # balanced by bases
            assert cover == list(range(n))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from minimap2_rs_b200 import shard
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(1)
    lens = rng.integers(5, 50, 101)
    offs = np.zeros(102, dtype=np.uint64)
    offs[1:] = np.cumsum(lens)
    cat = rng.integers(65, 70, int(offs[-1])).astype(np.uint8)
    names = ["q%d" % i for i in range(101)]

    def fake_map(c, o, nm):  # stands in for Context.map_batch(...).paf_lines(...): one line per read
        return ["%s\t%d\t%d" % (nm[i], int(o[i + 1] - o[i]), int(c[int(o[i]):int(o[i + 1])].sum())) for i in range(len(nm))]

    mine, (lo, hi) = shard.map_sharded(fake_map, cat, offs, names, world, rank)
    allv = shard.gather_lines(mine, dst=0)
    if rank == 0:
        q.put(allv == fake_map(cat, offs, names))
    else:
        assert allv is None
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_gather_in_input_order():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    ok = q.get(timeout=120)
    for p in ps:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok


def _plan_worker(rank, world, port, q):
    """every rank computes ITS share of the sharded index build (mm2_shard_plan, pure host code in libmm2b200.so); the shares
    are gathered over gloo and checked on rank 0"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import minimap2_rs_b200 as mm2
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lens = [193_750_000, 1, 193_750_000, 1, 50_000, 7, 0, 2038, 2039, 300_000_000]
    offs = np.zeros(len(lens) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(lens)
    res = {}
    for (w, k, flag) in ((10, 15, 0), (10, 19, 0), (5, 21, 0), (10, 14, 0), (10, 15, 1)):
        mine = mm2.shard_plan(offs, w, k, flag, world, rank)
        box = [None] * world
        dist.all_gather_object(box, mine)
        res[(w, k, flag)] = box
    if rank == 0:
        ok = True
        total = int(offs[-1])
        used = (total + 7) // 8
        for (w, k, flag), plans in res.items():
            tile = (k % 2 == 1) and (not flag or 9 <= w <= 64)   # even k -> whole sequences; HPC rides the w >= 9 tile kernel
            ok &= all(p["tile_path"] == int(tile) for p in plans)
            ok &= plans[0]["lo"] == 0 and all(plans[r]["hi"] == plans[r + 1]["lo"] for r in range(world - 1))   # shares tile the list
            if tile:
                T = 2048 - 2 * w if w >= 9 else 2048 - w   # positions (w >= 9) or steps per sketch tile
                ntiles = sum(max(1, (l + T - 1) // T) for l in lens)
                ok &= plans[-1]["hi"] == ntiles
                # a rank uploads about 1 / world of the genome although sequences are far from balanced
                ok &= all(abs(p["upload_bytes"] - total / world) < 0.02 * total for p in plans)
                for p in plans:   # halo: the sketch reads from 2w+k (+ alignment slack) before its first step
                    ok &= p["sketch_byte_lo"] % 16 == 0 and p["sketch_byte_hi"] <= total
            else:
                ok &= plans[-1]["hi"] == len(lens)
            ok &= plans[0]["word_lo"] == 0 and used <= plans[-1]["word_hi"] < used + world     # equal slots of ceil(used / world) words
            ok &= all(plans[r]["word_hi"] == plans[r + 1]["word_lo"] for r in range(world - 1))
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_sharded_build_plan():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_plan_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    ok = q.get(timeout=120)
    for p in ps:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok
