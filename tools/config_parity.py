"""GPU-vs-oracle comparisons of the BASELINE configs at their STATED sizes (tools/workloads.py), with timings.  Writes one JSON
per config under gpurun_out/ (copied to profiles/ when committed).

  python tools/config_parity.py --which c5 [--reads 50000] [--sample 1000]        # one GPU
  python tools/config_parity.py --which c4 [--reads 1000000] [--sample 10000]      # one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
         tools/config_parity.py --which c3 [--k 15,19]                            # sharded build, sha256 of the .mmi vs the CPU build

c3: every rank builds the 3.1 Gbp index with mm2_index_build_sharded; rank 0 writes the .mmi, the CPU oracle builds and writes
    its own, and the sha256 of both files (streamed) must be equal.  c4 / c5: the whole read set is mapped on the GPU; the first
    --sample reads are also aligned by the oracle on the full-size index and the PAF lines compared.
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import minimap2_rs_b200 as mm2  # noqa: E402
from tools import workloads as wl  # noqa: E402


def sha_file(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        while True:
            blk = f.read(1 << 24)
            if not blk:
                break
            h.update(blk)
    return h.hexdigest(), os.path.getsize(path)


def run_mapping(cfg, a):
    from oracle import orc
    ncpu = os.cpu_count() or 1
    w, k = wl.WK[cfg]
    sh = dict(wl.SHAPES[cfg])
    if a.reads:
        sh["reads"] = a.reads
    out = {"config": cfg, "reads": sh["reads"], "read_len": sh["read_len"], "k": k, "w": w}
    t0 = time.time()
    pins = []

    def pinned(n):
        pb = mm2.PinnedBuffer(n)
        pins.append(pb)
        return pb.array(np.uint8, n)

    if cfg == "c4":
        g, goffs, gnames = wl.genome_c3(alloc=pinned) if not a.genome_mbp else wl.genome_c3(16, int(a.genome_mbp * 1e6 / 16), alloc=pinned)
    else:
        g0, goffs, gnames = wl.genome_c5(int(a.genome_mbp * 1e6) if a.genome_mbp else wl.C1_LEN)
        g = pinned(g0.size)
        g[:] = g0
        del g0
    out["genome_bp"] = int(g.size)
    out["genome_gen_s"] = time.time() - t0
    ctx = mm2.Context(0)
    t0 = time.perf_counter()
    gi = mm2.Index.build(ctx, g, goffs, gnames, w=w, k=k)
    out["index_build_wall_s_first"] = time.perf_counter() - t0
    gi.close()
    t0 = time.perf_counter()
    gi = mm2.Index.build(ctx, g, goffs, gnames, w=w, k=k)
    out["index_build_wall_s_warm"] = time.perf_counter() - t0
    out["index_build_device_ms"] = gi.build_timings()
    out["index_stats"] = gi.stats()
    out["mid_occ"] = max(10, gi.calc_mid_occ())
    t0 = time.time()
    pr = pinned(sh["reads"] * sh["read_len"])
    _, roffs = wl.reads(cfg, g, goffs, sh["reads"], sh["read_len"], out=pr)
    out["reads_gen_s"] = time.time() - t0
    opts = mm2.default_map_opts(w, k)
    nwarm = min(sh["reads"], max(64, sh["reads"] // 50))
    ctx.map_batch(gi, pr[:int(roffs[nwarm])], roffs[:nwarm + 1], opts).close()      # warm-up: arenas, workers
    t0 = time.perf_counter()
    res = ctx.map_batch(gi, pr, roffs, opts)
    t_map = time.perf_counter() - t0
    out["map_wall_s"] = t_map
    out["mapped_bases_per_s"] = float(roffs[-1]) / t_map
    out["stage_ms_summed_over_subbatches"] = {kk: round(v, 2) for kk, v in ctx.last_timings().items()}
    out["stats"] = dict(res.stats)
    out["paf_records"] = int(res.n_recs)
    names = ["r%07d" % i for i in range(sh["reads"])]
    nsample = min(a.sample, sh["reads"])
    got = [l for l in res.paf_lines(names) if int(l.split("\t")[0][1:]) < nsample]
    res.close()
    # DP cells of a slice of the batch (device counter; an untimed pass)
    ncell = min(sh["reads"], a.cell_reads)
    ctx.count_cells(True)
    t0 = time.perf_counter()
    r2 = ctx.map_batch(gi, pr[:int(roffs[ncell])], roffs[:ncell + 1], opts)
    t_cells = time.perf_counter() - t0
    cells = ctx.last_cells
    out["cell_pass"] = {"reads": ncell, "cells": int(cells), "anchors": int(r2.stats["n_anchors"]), "cells_per_anchor": cells / max(1, r2.stats["n_anchors"]),
                        "wall_s_with_counter": t_cells, "chain_ms": ctx.last_timings().get("chain")}
    r2.close()
    ctx.count_cells(False)
    # whole-job rate: DP cells of the batch (extrapolated from the counted slice) over the WALL time of the mapping call (the stage
    # timers are summed over concurrent worker contexts and would double-count)
    est_cells = cells / max(1, ncell) * sh["reads"]
    out["chain"] = {"estimated_cells_whole_batch": est_cells, "gcells_per_s_over_map_wall": est_cells / 1e9 / t_map,
                    "frac_of_int_peak_20_ops_per_cell_over_map_wall": 20 * est_cells / t_map / (148 * 128 * 1.965e9),
                    "counted_slice_gcells_per_s": cells / 1e9 / max(1e-9, t_cells)}
    # the oracle on the sample, full-size index
    t0 = time.time()
    oi = orc.Index.build(g, goffs, gnames, w=w, k=k, threads=ncpu)
    out["oracle_index_build_s"] = time.time() - t0
    out["index_stats_equal_oracle"] = bool(gi.stats() == oi.stats() and gi.calc_mid_occ() == oi.calc_mid_occ())
    t0 = time.time()
    want, st = oi.align_batch(np.array(pr[:int(roffs[nsample])]), roffs[:nsample + 1], names[:nsample], orc.AlignOpts.default(w, k), threads=ncpu)
    out["oracle_sample"] = {"reads": nsample, "threads": ncpu, "seconds": time.time() - t0, "bases_per_s": float(roffs[nsample]) / max(1e-9, st.seconds),
                            "cells_per_anchor": st.cells / max(1, st.n_anchors), "anchors_per_read": st.n_anchors / max(1, nsample), "rescued": int(st.n_rescued)}
    out["paf_identical_on_sample"] = bool(got == want)
    if got != want:
        bad = [i for i, (x, y) in enumerate(zip(got, want)) if x != y][:3]
        out["first_mismatches"] = [(got[i], want[i]) for i in bad] + [len(got), len(want)]
    return out


def run_c3(a):
    import torch
    import torch.distributed as dist
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ctx = mm2.Context(lr)
    comm = mm2.Comm.from_torch(ctx, dist) if world > 1 else None
    pins = []

    def pinned(n):
        pb = mm2.PinnedBuffer(n)
        pins.append(pb)
        return pb.array(np.uint8, n)

    t0 = time.time()
    g, goffs, gnames = wl.genome_c3(alloc=pinned) if not a.genome_mbp else wl.genome_c3(16, int(a.genome_mbp * 1e6 / 16), alloc=pinned)
    out = {"config": "c3", "n_gpus": world, "genome_bp": int(g.size), "records": len(gnames), "genome_gen_s": time.time() - t0, "builds": []}
    for k in [int(x) for x in a.k.split(",")]:
        e = {"k": k, "w": 10}
        times = []
        for rep in range(3):
            torch.cuda.synchronize()
            if comm:
                comm.barrier()
            t0 = time.perf_counter()
            gi = mm2.Index.build_sharded(ctx, comm, g, goffs, gnames, w=10, k=k) if comm else mm2.Index.build(ctx, g, goffs, gnames, w=10, k=k)
            torch.cuda.synchronize()
            if comm:
                comm.barrier()
            times.append(time.perf_counter() - t0)
            if rep < 2:
                gi.close()
        t = torch.tensor([min(times)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e["sharded_build_s"] = float(t[0])
        e["gbp_per_s"] = g.size / float(t[0]) / 1e9
        e["device_ms_rank0"] = gi.build_timings()
        e["stages_ms_rank0"] = {kk: round(v, 2) for kk, v in ctx.last_timings().items()}
        e["stats"] = gi.stats()
        if rank == 0:
            t0 = time.perf_counter()
            g1 = mm2.Index.build(ctx, g, goffs, gnames, w=10, k=k)
            g1.close()
            t0 = time.perf_counter()
            g1 = mm2.Index.build(ctx, g, goffs, gnames, w=10, k=k)
            e["one_gpu_build_s"] = time.perf_counter() - t0
            e["one_gpu_device_ms"] = g1.build_timings()
            e["speedup_vs_one_gpu"] = e["one_gpu_build_s"] / e["sharded_build_s"]
            g1.close()
            if not a.no_sha:
                pg = "/tmp/c3_gpu_k%d.mmi" % k
                t0 = time.time()
                gi.save_to_mmi(pg)
                e["gpu_mmi_sha256"], e["mmi_bytes"] = sha_file(pg)
                e["save_and_sha_s"] = time.time() - t0
                os.remove(pg)
            if not a.no_cpu:
                from oracle import orc
                t0 = time.time()
                oi = orc.Index.build(g, goffs, gnames, w=10, k=k, threads=os.cpu_count() or 1)
                e["oracle_build_s"] = time.time() - t0
                e["oracle_threads"] = os.cpu_count() or 1
                po = "/tmp/c3_cpu_k%d.mmi" % k
                oi.save_mmi(po)
                e["cpu_mmi_sha256"], _ = sha_file(po)
                os.remove(po)
                e["mmi_byte_identical_to_cpu"] = bool(e["gpu_mmi_sha256"] == e["cpu_mmi_sha256"])
                e["stats_equal_cpu"] = bool(gi.stats() == oi.stats() and gi.calc_mid_occ() == oi.calc_mid_occ())
                oi.close()
        gi.close()
        if world > 1:
            dist.barrier()
        out["builds"].append(e)
    if world > 1:
        dist.destroy_process_group()
    return out if rank == 0 else None


def run_c3cpu(a):
    """the CPU side of configs[2]: the oracle builds the 3.1 Gbp index (k = 15, 19), writes the .mmi and hashes it; the sharded GPU
    build of a `--which c3 --no-cpu` run on the same seeded genome must produce the same sha256"""
    from oracle import orc
    ncpu = os.cpu_count() or 1
    g, goffs, gnames = wl.genome_c3() if not a.genome_mbp else wl.genome_c3(16, int(a.genome_mbp * 1e6 / 16))
    out = {"config": "c3cpu", "genome_bp": int(g.size), "records": len(gnames), "threads": ncpu, "builds": []}
    for k in [int(x) for x in a.k.split(",")]:
        t0 = time.time()
        oi = orc.Index.build(g, goffs, gnames, w=10, k=k, threads=ncpu)
        e = {"k": k, "w": 10, "oracle_build_s": time.time() - t0, "gbp_per_s": g.size / (time.time() - t0) / 1e9, "stats": oi.stats(), "mid_occ": oi.calc_mid_occ()}
        po = "/tmp/c3_cpu_k%d.mmi" % k
        t0 = time.time()
        oi.save_mmi(po)
        e["cpu_mmi_sha256"], e["mmi_bytes"] = sha_file(po)
        e["save_and_sha_s"] = time.time() - t0
        os.remove(po)
        oi.close()
        out["builds"].append(e)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", required=True, choices=["c3", "c3cpu", "c4", "c5"])
    ap.add_argument("--reads", type=int, default=0)
    ap.add_argument("--sample", type=int, default=1000)
    ap.add_argument("--cell-reads", type=int, default=2000, help="reads of the DP-cell counting pass")
    ap.add_argument("--genome-mbp", type=float, default=0.0)
    ap.add_argument("--k", default="15,19")
    ap.add_argument("--no-cpu", action="store_true", help="c3: skip the CPU build (its sha256 comes from a `--which c3cpu` run on the same genome)")
    ap.add_argument("--no-sha", action="store_true")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    out = run_c3(a) if a.which == "c3" else run_c3cpu(a) if a.which == "c3cpu" else run_mapping(a.which, a)
    if out is not None:
        path = a.out or os.path.join(ROOT, "gpurun_out", "config_%s.json" % a.which)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        json.dump(out, open(path, "w"), indent=1)
        print(json.dumps(out))


if __name__ == "__main__":
    main()
