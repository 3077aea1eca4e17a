#!/usr/bin/env python
"""bench.py — mapped bases/s of the B200 mapping hot path (sketch -> filter -> lookup -> anchor sort -> chain -> PAF record)
and index-build Gbp/s, on the BASELINE.json configs (tools/workloads.py).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c2|c4|c5|c3] [--reads R] [--genome-mbp G]

  c2 (default)  configs[1]: 145 Mbp genome (k=15 w=10) + 100k x 10 kb ONT-like reads PER GPU per step   (weak scaling)
  c4            configs[3]: 3.1 Gbp genome (k=19 w=10) + 1 M x 15 kb HiFi-like reads, sharded over the GPUs (strong scaling)
  c5            configs[4]: 145 Mbp repeat-rich genome + 50k x 100 kb reads at 8 % error, sharded            (strong scaling)
  c3            configs[2]: index build of the 3.1 Gbp genome, bucket-sharded over the GPUs; metric = index-build Gbp/s

A mapping step = one pass of the whole mapping path over one batch of reads.  `value` is measured with the reads resident
in HBM (CUDA events on the launching stream); `e2e` goes through mm2_map_batch with pinned HOST buffers (H2D of the ASCII
reads, D2H of the records and the host-side record assembly inside the timed region, wall clock).  Under torchrun every
rank maps its own shard of reads against its own replica of the index (no collective on the data path), and the index of
the run is ALSO built once more by all ranks together (bucket-sharded build over NCCL inside libmm2b200.so), timed, and
reported as `index_build.sharded`.  `--impl reference` times the CPU restatement of the reference (oracle/, a C++ port:
the Rust crate cannot be built in this image) with all host threads on a bounded sample.
"""
import argparse
import glob
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# NCCL_DEBUG is left to the caller (the driver reads NCCL's own log); the JSON line is the LAST line of stdout

INT_PEAK = 148 * 128 * 1.965e9   # 148 SMs x 128 INT32 lanes x 1.965 GHz = 37.2 Tint-op/s (nominal)
OPS_PER_CELL = 20                # integer operations credited per DP cell (lchain.rs:80 iteration)
LIMITER = {"sketch": "instruction issue (ncu: issue-active well above DRAM %)", "chain": "integer pipe / warp issue",
           "lookup": "DRAM on random 16-byte probes (a 64-byte access each)", "anchor_sort": "shared memory / issue"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per unit of each stage's dominant kernel, from the newest profiles/r*_ncu_traffic.json (written by
    tools/ncu_traffic.py from an `ncu --set full` capture, with the commit it was taken at)"""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_traffic.json")))
    if not files:
        return {}, None
    d = json.load(open(files[-1]))
    return d.get("kernels", {}), {"file": os.path.relpath(files[-1], ROOT), "commit": d.get("commit"), "workload": d.get("workload")}


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---- workloads -------------------------------------------------------------------------------------------------------------
def shape(args):
    from tools import workloads as wl
    sh = dict(wl.SHAPES[args.config]) if args.config in wl.SHAPES else dict(reads=0, read_len=0)
    if args.reads:
        sh["reads"] = args.reads
    if args.read_len:
        sh["read_len"] = args.read_len
    return sh


def make_genome(args, alloc=None):
    from tools import workloads as wl
    if args.config in ("c2", "c1"):
        return wl.genome_c2(int(args.genome_mbp * 1e6) if args.genome_mbp else wl.C1_LEN)
    if args.config == "c5":
        return wl.genome_c5(int(args.genome_mbp * 1e6) if args.genome_mbp else wl.C1_LEN)
    if args.genome_mbp:   # reduced C3/C4 genome: 16 chromosomes of genome_mbp / 16 each
        return wl.genome_c3(16, int(args.genome_mbp * 1e6 / 16), alloc=alloc)
    return wl.genome_c3(alloc=alloc)


def make_reads(args, g, goffs, rank, world, out=None):
    """this rank's reads: c2 = its own read set (weak scaling); c4 / c5 = its contiguous shard of the one read set (strong)"""
    from tools import gen, workloads as wl
    sh = shape(args)
    if args.config == "c2":
        return gen.reads(wl.SEED_C2_READS + rank, g, goffs, sh["reads"], sh["read_len"], *wl.ERR["c2"], out=out)
    lo, hi = sh["reads"] * rank // world, sh["reads"] * (rank + 1) // world
    return wl.reads(args.config, g, goffs, hi - lo, sh["read_len"], first=lo, out=out)


def workload_config(args, world):
    from tools import workloads as wl
    sh = shape(args)
    w, k = wl.WK[args.config]
    desc = {"c2": "BASELINE configs[1]: synthetic %.1f Mbp random genome, index k=%d w=%d b=14; %d simulated %d bp ONT-like reads (10 %% error) per GPU per step",
            "c4": "BASELINE configs[3]: synthetic %.1f Mbp genome (16 chromosomes at even rids + 1-bp N records), index k=%d w=%d b=14; %d simulated %d bp HiFi-like reads (0.5 %% error) per step over all GPUs",
            "c5": "BASELINE configs[4]: synthetic %.1f Mbp repeat-rich genome (40 %% tandem arrays, 20 %% dispersed families), index k=%d w=%d b=14; %d simulated %d bp reads (8 %% error) per step over all GPUs",
            "c3": "BASELINE configs[2]: index build of a synthetic %.1f Mbp genome (16 chromosomes at even rids + 1-bp N records, 0.1 %% N), k=%d w=%d b=14%.0s%.0s"}[args.config]
    glen = args.genome_mbp or (3100.000015 if args.config in ("c3", "c4") else 145.138636)
    cfg = {"workload": desc % (glen, k, w, sh["reads"], sh["read_len"]), "name": args.config,
           "parallelism": ("reads sharded over %d GPU(s), index replicated" if args.config != "c3" else "minimizers sharded by hash-prefix bucket over %d GPU(s), index replicated") % world}
    if args.config != "c3":
        cfg.update(reads=sh["reads"], read_len=sh["read_len"],
                   l2="inputs (%.2f GB of reads per GPU per step) are larger than the 126 MB L2" % (sh["reads"] * sh["read_len"] / (1 if args.config == "c2" else world) / 1e9))
    else:
        cfg["l2"] = "inputs (the genome and its minimizers, GBs) are larger than the 126 MB L2"
    return cfg


# ---- the CPU arm ---------------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """CPU restatement of the reference on the host cores (rank 0 only), bounded sample of the same workload"""
    if rank != 0:
        return
    from oracle import orc
    from tools import workloads as wl
    ncpu = os.cpu_count() or 1
    w, k = wl.WK[args.config]
    sh = shape(args)
    if args.config == "c3":
        # index build: the CPU port builds a BOUNDED sample of the 3.1 Gbp genome (4 of its 16 chromosomes by default)
        chroms = args.ref_chroms
        g, goffs, names = wl.genome_c3(chroms) if not args.genome_mbp else wl.genome_c3(chroms, int(args.genome_mbp * 1e6 / 16))
        times = []
        for it in range(args.warmup + args.steps):
            t0 = time.time()
            oi = orc.Index.build(g, goffs, names, w=w, k=k, threads=ncpu)
            dt = time.time() - t0
            oi.close()
            if it >= args.warmup:
                times.append(dt)
        ms = 1e3 * float(np.mean(times))
        val = g.size / (ms / 1e3)
        line = {"impl": "reference", "metric": "index_build_bases_per_sec", "value": val, "unit": "bases/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
                "data": "synthetic", "config": workload_config(args, world),
                "cpu_baseline": {"value": val, "unit": "bases/s", "cores": ncpu, "kind": "port",
                                 "sample": "%d of the 16 chromosomes (%.2f Gbp) per step, %d host threads (rayon-style: sketch over sequences, post-process over buckets)" % (chroms, g.size / 1e9, ncpu)},
                "e2e": {"value": val, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return
    g, goffs, gnames = make_genome(args)
    t0 = time.time()
    oi = orc.Index.build(g, goffs, gnames, w=w, k=k, threads=ncpu)
    t_build = time.time() - t0
    sample = min(sh["reads"], args.ref_sample if args.config == "c2" else {"c4": 12_000, "c5": 48}[args.config])
    cat, roffs = make_reads(args, g, goffs, 0, max(1, sh["reads"] // sample) if args.config != "c2" else 1)
    sample = min(sample, roffs.size - 1)
    names = ["r%06d" % i for i in range(sample)]
    opts = orc.AlignOpts.default(w, k)
    times = []
    bases = int(roffs[sample])
    for it in range(args.warmup + args.steps):
        _, st = oi.align_batch(cat[:bases], roffs[:sample + 1], names, opts, threads=ncpu)
        if it >= args.warmup:
            times.append(st.seconds)
    ms = 1e3 * float(np.mean(times))
    val = bases / (ms / 1e3)
    line = {
        "impl": "reference", "metric": "mapped_bases_per_sec", "value": val, "unit": "bases/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak" if args.config == "c2" else "strong", "vs_baseline": None,
        "dtype": "u64/i32", "data": "synthetic", "config": workload_config(args, world),
        "cpu_baseline": {"value": val, "unit": "bases/s", "cores": ncpu, "kind": "port",
                         "what": "oracle/ (C++ restatement of mm2rs; the Rust crate cannot be built in this image), reads split over all host threads",
                         "sample": "%d reads x %d bp per step (of %d), %d host threads; index built by the same CPU port in %.1f s (%.4f Gbp/s)"
                                   % (sample, sh["read_len"], sh["reads"], ncpu, t_build, g.size / t_build / 1e9)},
        "e2e": {"value": val, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "index_build": {"seconds": t_build, "gbp_per_s": g.size / t_build / 1e9, "threads": ncpu},
    }
    print(json.dumps(line))


# ---- the GPU arm ---------------------------------------------------------------------------------------------------------------
def timed_builds(mm2, ctx, pg, goffs, gnames, w, k, comm, reps=3):
    """-> (index, wall seconds of the first build, best warm wall seconds, device-ms breakdown of the last build)"""
    import torch

    def build():
        if comm is not None:
            return mm2.Index.build_sharded(ctx, comm, pg, goffs, gnames, w=w, k=k, b=14)
        return mm2.Index.build(ctx, pg, goffs, gnames, w=w, k=k, b=14)

    def sync():
        torch.cuda.synchronize()
        if comm is not None:
            comm.barrier()

    sync()
    t0 = time.perf_counter()
    gi = build()
    sync()
    t_first = time.perf_counter() - t0
    build().close()   # warm-up: kernels loaded, arenas and the index pool grown
    best, bt = 1e9, None
    for _ in range(reps):
        sync()
        t0 = time.perf_counter()
        g2 = build()   # from pinned host ASCII to an index resident in HBM (on every rank when sharded)
        sync()
        best = min(best, time.perf_counter() - t0)
        bt = g2.build_timings()
        g2.close()
    return gi, t_first, best, bt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--reads", type=int, default=0, help="override the config's read count (reduced runs)")
    ap.add_argument("--read-len", type=int, default=0)
    ap.add_argument("--genome-mbp", type=float, default=0.0, help="override the config's genome size (reduced runs)")
    ap.add_argument("--ref-sample", type=int, default=20_000, help="reads per step of the CPU reference arm (c2)")
    ap.add_argument("--ref-chroms", type=int, default=4, help="chromosomes per step of the CPU reference arm (c3)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads of the single-thread cpu_baseline leg (default per config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import minimap2_rs_b200 as mm2
    from tools import workloads as wl
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    w, k = wl.WK[args.config]
    stream = torch.cuda.current_stream()
    ctx = mm2.Context(local_rank, stream=stream.cuda_stream)
    comm = mm2.Comm.from_torch(ctx, dist) if world > 1 else None   # NCCL communicator owned by libmm2b200 (id exchanged through torch)

    # ---- genome in pinned host memory; index build (single GPU and, under torchrun, bucket-sharded over all ranks) -----------------------
    pins = []

    def pinned(n):
        pb = mm2.PinnedBuffer(n, device=local_rank)
        pins.append(pb)
        return pb.array(np.uint8, n)

    t0 = time.time()
    if args.config in ("c3", "c4"):
        pg, goffs, gnames = make_genome(args, alloc=pinned)
    else:
        g, goffs, gnames = make_genome(args)
        pg = pinned(g.size)
        pg[:] = g
        del g
    t_gen = time.time() - t0
    gi, t_first, t_warm, bt = timed_builds(mm2, ctx, pg, goffs, gnames, w, k, None, reps=2 if args.config in ("c3", "c4") else 3)
    n_keys = gi.stats()[0]
    index_build = {"genome_bp": int(pg.size), "wall_s_first": t_first, "wall_s_warm": t_warm, "gbp_per_s_warm": pg.size / t_warm / 1e9,
                   "device_ms": bt, "n_keys": int(n_keys), "what": "one GPU: pinned host ASCII -> index resident in HBM"}
    if comm is not None:
        sig = (gi.stats(), gi.calc_mid_occ())
        big = args.config in ("c3", "c4")
        if big:        # two replicas of a 3.1 Gbp index (2 x 45 GB) plus the build arenas do not fit next to each other
            gi.close()
        gs, ts_first, ts_warm, bts = timed_builds(mm2, ctx, pg, goffs, gnames, w, k, comm, reps=2 if big else 3)
        tt = torch.tensor([ts_warm, t_warm], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        same = (gs.stats(), gs.calc_mid_occ()) == sig
        index_build["sharded"] = {"n_gpus": world, "wall_s_warm_max_over_ranks": float(tt[0]), "gbp_per_s": pg.size / float(tt[0]) / 1e9,
                                  "speedup_vs_one_gpu": float(tt[1]) / float(tt[0]), "device_ms": bts, "stats_equal_single_gpu_build": bool(same),
                                  "what": "bucket-sharded build over NCCL inside libmm2b200 (mm2_index_build_sharded): genome in host memory -> replicated index on every GPU"}
        if big:
            gi = gs    # map against the replica the sharded build left on this GPU
        else:
            gs.close()

    if args.config == "c3":
        run_c3(args, mm2, ctx, comm, pg, goffs, gnames, gi, index_build, rank, world, local_rank, barrier)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- reads: pinned host copy (e2e) and HBM-resident copy (value) ---------------------------------------------------------
    sh = shape(args)
    nr_rank = sh["reads"] if args.config == "c2" else (sh["reads"] * (rank + 1) // world - sh["reads"] * rank // world)
    pr = pinned(nr_rank * sh["read_len"])
    _, roffs = make_reads(args, pg, goffs, rank, world, out=pr)
    n_bases = int(roffs[-1])
    resident_ok = n_bases <= (48 << 30)
    d_cat = torch.empty(n_bases + 64, dtype=torch.uint8, device="cuda")
    d_cat[:n_bases].copy_(torch.from_numpy(pr))
    d_off = torch.from_numpy(roffs.astype(np.int64)).cuda()
    opts = mm2.default_map_opts(w, k)
    torch.cuda.synchronize()

    def step_dev():
        return ctx.map_batch(gi, None, roffs, opts, device_ptrs=(d_cat.data_ptr(), d_off.data_ptr()))

    def step_e2e():
        return ctx.map_batch(gi, pr, roffs, opts)

    steps = args.steps
    for _ in range(args.warmup):
        step_dev().close()
    # DP cells of one step: one untimed pass with the cell counter on (the counter costs kernel time; cells do not change between steps)
    ctx.count_cells(True)
    step_dev().close()
    cells = ctx.last_cells
    ctx.count_cells(False)
    step_dev().close()
    # ---- timed region: device-resident -----------------------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.launch_count
    stage_ms = {}
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        res = step_dev()
        for kname, v in ctx.last_timings().items():
            stage_ms[kname] = stage_ms.get(kname, 0.0) + v / steps
        stats = dict(res.stats)
        n_recs = res.n_recs
        res.close()
    ev1.record(stream)
    barrier()
    dev_ms = ev0.elapsed_time(ev1) / steps
    launches = (ctx.launch_count - l0) // steps
    # ---- timed region: end to end from pinned host memory ------------------------------------------------------------------------
    step_e2e().close()
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(steps):
        res = step_e2e()
        d2h = nr_rank * 64 + 8
        res.close()
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / steps
    barrier()
    # ---- the same from pinned host memory as 2-bit codes (mm2_map_batch_packed): a separate number, `e2e` stays the ASCII call ----
    e2e_packed = None
    if args.config == "c2":
        pk = pinned(n_bases // 4 + 128)
        t0 = time.perf_counter()
        _, n_pos = mm2.pack_reads(pr, out=pk)
        pack_s = time.perf_counter() - t0
        ctx.map_batch_packed(gi, pk, n_pos, roffs, opts).close()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            res = ctx.map_batch_packed(gi, pk, n_pos, roffs, opts)
            n_recs_pk = res.n_recs
            res.close()
        torch.cuda.synchronize()
        e2e_packed = [1e3 * (time.perf_counter() - t0) / steps, pack_s, int(n_bases // 4 + n_pos.size * 8 + roffs.size * 8), n_recs_pk]
        barrier()
    clocks = sampler.stop()

    # ---- max over ranks ------------------------------------------------------------------------------------------------------------
    t = torch.tensor([dev_ms, e2e_ms, e2e_packed[0] if e2e_packed else 0.0], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(n_bases), float(cells), float(stats["n_anchors"]), float(stats["n_minimizers"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms, e2e_pk_ms = float(t[0]), float(t[1]), float(t[2])
    total_bases = float(tot[0])

    if rank == 0:
        peak, peak_src = peaks()
        traffic, traffic_src = ncu_traffic()
        nm, na = stats["n_minimizers"], stats["n_anchors"]
        # algorithmic work per launch of rank 0 (SURVEY.md §8d / DESIGN.md §3): sketch L + 16 n_min; lookup 16 n_min + 16 n_min;
        # anchors 8 per occurrence + sort 32 n_anchor; chain 20 integer ops per DP cell
        alg = {"sketch": n_bases + 16 * nm, "lookup": 32 * nm, "anchor_sort": 8 * na + 32 * na}
        units = {"base": n_bases, "anchor": na, "minimizer": nm}
        kernels = {}
        dev_stages = {kn: ms for kn, ms in stage_ms.items() if not kn.startswith("host_") and kn not in ("h2d", "d2h", "end")}
        for kname, ms in stage_ms.items():
            e = {"ms": ms}
            if kname in dev_stages:
                e["share"] = ms / max(1e-9, sum(dev_stages.values()))
            if kname in alg and ms > 0:
                e.update(bound="hbm", algorithmic_GB=alg[kname] / 1e9, achieved_GBps=alg[kname] / 1e9 / (ms / 1e3))
                e["frac_of_hbm_peak"] = e["achieved_GBps"] / peak
            if kname == "chain" and ms > 0:
                e.update(bound="int", cells=int(cells), cells_per_anchor=cells / max(1, na), gcells_per_s=cells / 1e9 / (ms / 1e3),
                         int_ops=OPS_PER_CELL * int(cells), achieved_Tintops=OPS_PER_CELL * cells / 1e12 / (ms / 1e3))
                e["frac_of_int_peak"] = e["achieved_Tintops"] * 1e12 / INT_PEAK
            if kname in traffic and ms > 0:
                tr = traffic[kname]
                e["dram_traffic_GB_ncu"] = tr["dram_bytes"] / tr["units"] * units[tr["unit"]] / 1e9
            if kname in LIMITER:
                e["limiter"] = LIMITER[kname]
            kernels[kname] = e
        dom = max(dev_stages, key=dev_stages.get)
        tr = traffic.get(dom)
        tr_bytes = tr["dram_bytes"] / tr["units"] * units[tr["unit"]] if tr else None
        if dom == "chain":
            roof = {"kernel": "chain (chain_ring_kernel / chain_dense_kernel)", "bound": "int", "achieved": kernels[dom]["achieved_Tintops"], "peak": INT_PEAK / 1e12,
                    "unit": "Tint-op/s", "frac": kernels[dom]["frac_of_int_peak"], "traffic": tr_bytes, "peak_source": "nominal: 148 SMs x 128 INT32 lanes x 1.965 GHz (SURVEY.md 8d)",
                    "note": "20 integer ops per DP cell (lchain.rs:80 iteration), cells counted on the device in an untimed pass of the same step; %.1f Gcells/s" % kernels[dom]["gcells_per_s"]}
        else:
            roof = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["achieved_GBps"], "peak": peak, "unit": "GB/s",
                    "frac": kernels[dom]["achieved_GBps"] / peak, "traffic": tr_bytes, "peak_source": peak_src,
                    "note": "limiter: " + LIMITER.get(dom, "?") + "; algorithmic bytes per SURVEY.md 8d"}
        roof["traffic_source"] = traffic_src
        line = {
            "metric": "mapped_bases_per_sec", "value": total_bases / (dev_ms / 1e3), "unit": "bases/s", "n_gpus": world, "steps": steps,
            "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak" if args.config == "c2" else "strong", "vs_baseline": None,
            "dtype": "u64/i32", "data": "synthetic", "config": workload_config(args, world),
            "e2e": {"value": total_bases / (e2e_ms / 1e3), "unit": "bases/s", "h2d_bytes_per_step": int(n_bases + roffs.size * 8),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms, "timing": "wall clock around mm2_map_batch (pinned host ASCII reads in, PAF records out)"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "kernels": kernels, "index_build": index_build,
            "work": {"reads": int(sh["reads"] * (world if args.config == "c2" else 1)), "bases_per_step": total_bases, "minimizers": int(tot[3]), "anchors": int(tot[2]),
                     "dp_cells": int(tot[1]), "paf_records_rank0": n_recs, "rescued_rank0": int(stats["n_rescued"]), "minimizers_kept_rank0": int(stats["n_minimizers_kept"])},
        }
        if e2e_packed:
            line["e2e_packed"] = {"value": total_bases / (e2e_pk_ms / 1e3), "unit": "bases/s", "ms_per_step": e2e_pk_ms, "h2d_bytes_per_step": e2e_packed[2],
                                  "paf_records_rank0": e2e_packed[3], "same_record_count_as_e2e": e2e_packed[3] == n_recs,
                                  "host_pack_s_once_untimed": e2e_packed[1],
                                  "what": "wall clock around mm2_map_batch_packed: the same reads as 2-bit codes + N positions in pinned host memory (0.25 B per base over the link); packing (mm2_pack_reads, one host thread) is done once outside the timed region, as for a caller that already holds packed reads; `e2e` above is the drop-in ASCII call"}
        if not args.no_cpu_baseline and world == 1:   # the CPU leg is an N = 1 measurement (rank 0's host cores are shared with the other ranks otherwise)
            from oracle import orc
            ncpu = os.cpu_count() or 1
            t0 = time.time()
            oi = orc.Index.build(pg, goffs, gnames, w=w, k=k, threads=ncpu)
            t_ob = time.time() - t0
            sample = min(nr_rank, args.cpu_sample or {"c2": 3000, "c4": 1500, "c5": 6}[args.config])
            names = ["r%06d" % i for i in range(sample)]
            sub = np.array(pr[:int(roffs[sample])])
            lines_cpu, st = oi.align_batch(sub, roffs[:sample + 1], names, orc.AlignOpts.default(w, k), threads=1)
            res = ctx.map_batch(gi, sub, roffs[:sample + 1], opts)
            same = res.paf_lines(names) == lines_cpu
            res.close()
            line["cpu_baseline"] = {"value": int(roffs[sample]) / st.seconds, "unit": "bases/s", "cores": 1, "kind": "port",
                                    "what": "oracle/ (C++ restatement of mm2rs; the Rust crate cannot be built in this image)",
                                    "sample": "first %d reads of the step on ONE host thread (what `mm2rs align` does per read); "
                                              "CPU index build with %d threads took %.1f s (%.3f Gbp/s)" % (sample, ncpu, t_ob, pg.size / t_ob / 1e9),
                                    "paf_identical_to_gpu_on_sample": bool(same), "cells_per_anchor": st.cells / max(1, st.n_anchors)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_c3(args, mm2, ctx, comm, pg, goffs, gnames, gi, index_build, rank, world, local_rank, barrier):
    """configs[2]: the metric is index-build bases/s; a step = one build of the whole genome (sharded over the ranks when N > 1)"""
    import torch
    import torch.distributed as dist
    from tools import workloads as wl
    w, k = wl.WK["c3"]
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.launch_count
    times, bt = [], None
    for it in range(args.warmup + args.steps):
        barrier()
        t0 = time.perf_counter()
        g2 = mm2.Index.build_sharded(ctx, comm, pg, goffs, gnames, w=w, k=k, b=14) if comm is not None else mm2.Index.build(ctx, pg, goffs, gnames, w=w, k=k, b=14)
        barrier()
        if it >= args.warmup:
            times.append(time.perf_counter() - t0)
            bt = g2.build_timings()
        if it == args.warmup:
            l1 = ctx.launch_count
        g2.close()
    launches = ctx.launch_count - l1 if args.steps > 1 else l1 - l0
    clocks = sampler.stop()
    t = torch.tensor([float(np.mean(times))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t[0])
    if rank == 0:
        peak, peak_src = peaks()
        nmin = bt["n_minimizers"]
        sort_ms = bt["sort_ms"]
        sort_bytes = 32 * nmin / world
        line = {"metric": "index_build_bases_per_sec", "value": pg.size / sec, "unit": "bases/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * sec, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": workload_config(args, world),
                "e2e": {"value": pg.size / sec, "unit": "bases/s", "h2d_bytes_per_step": int(pg.size / world), "d2h_bytes_per_step": 0, "ms_per_step": 1e3 * sec,
                        "timing": "wall clock around the build call: genome ASCII in pinned host memory -> index resident in HBM on every rank (value and e2e are the same measurement: the build starts from host memory by definition)"},
                "gpu_launches": int(max(1, launches // max(1, args.steps - 1 if args.steps > 1 else 1))), "clocks": clocks,
                "roofline": {"kernel": "radix sort of (key, position) pairs", "bound": "hbm", "achieved": sort_bytes / 1e9 / (sort_ms / 1e3) if sort_ms > 0 else None, "peak": peak,
                             "unit": "GB/s", "frac": (sort_bytes / 1e9 / (sort_ms / 1e3) / peak) if sort_ms > 0 else None, "traffic": None, "peak_source": peak_src,
                             "note": "algorithmic bytes = 32 per minimizer (one read + one write of each 16-byte record, SURVEY.md 8d); rank 0's share"},
                "index_build": index_build, "device_ms": bt}
        print(json.dumps(line))


if __name__ == "__main__":
    main()
