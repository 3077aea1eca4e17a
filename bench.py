#!/usr/bin/env python
"""bench.py — mapped bases/s of the B200 mapping hot path (sketch -> filter -> lookup -> anchor sort -> chain -> PAF record)
on BASELINE.json configs[1]: synthetic 145 Mbp genome (k=15, w=10) + 100k simulated 10 kb ONT-like reads (~10 % error).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--reads R] [--genome-mbp G]

A step = one pass of the whole mapping path over one batch of reads.  `value` is measured with the reads resident in
HBM (CUDA events on the launching stream); `e2e` goes through mm2_map_batch with pinned HOST buffers (H2D of the reads,
D2H of the records and the host-side record assembly inside the timed region, wall clock).  Under torchrun every rank
maps its own shard of reads against its own replica of the index (no collective on the data path; weak scaling).
`--impl reference` times the CPU restatement of the reference (oracle/) with all host threads on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line (NCCL_DEBUG=VERSION prints a banner)

GENOME_SEED, READS_SEED = 0xB2000002, 0xB2001002
W, K = 10, 15
ERR = (0.0333, 0.0333, 0.0333)  # sub / ins / del
# dram__bytes_read.sum + dram__bytes_write.sum of one launch in the `ncu --set full` capture of tools/prof_step.py
# (profiles/r01_ncu_v11.md): (bytes, units that launch processed, unit)
NCU_TRAFFIC = {"sketch": (0.1011e9 + 0.2471e9, 100e6, "base"), "chain": (0.1393e9 + 0.1882e9, 6.978e6, "anchor"),
               "lookup": (1.2007e9 + 0.0907e9, 18.628e6, "minimizer"), "anchor_sort": (0.1637e9 + 0.0725e9, 6.978e6, "anchor")}
NCU_NOTE = {"sketch": "sketch_tile_kernel_v3 is instruction-issue bound (ncu: 63 % issue-active at 56 % occupancy, 5 % DRAM throughput), not HBM bound",
            "chain": "chain_ring_kernel is warp-issue bound (ncu: 76 % issue-active at 40 % occupancy, 5 % DRAM throughput), not HBM bound",
            "lookup": "seed_hits_kernel is DRAM bound on random 16-byte probes that cost a 64-byte access each (ncu: 48 % DRAM throughput)",
            "anchor_sort": "anchor_msort_kernel is issue / shared-memory bound (ncu: 60 % issue-active)"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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


def make_workload(args, rank):
    from tools import gen
    glen = int(args.genome_mbp * 1e6) if args.genome_mbp else 145_138_636
    t0 = time.time()
    g = gen.genome(GENOME_SEED, glen)
    goffs = np.array([0, glen], dtype=np.uint64)
    cat, roffs = gen.reads(READS_SEED + rank, g, goffs, args.reads, args.read_len, *ERR)
    return g, goffs, cat, roffs, time.time() - t0


def run_reference(args, rank, world):
    """CPU restatement of the reference on the host cores (rank 0 only)."""
    if rank != 0:
        return
    from oracle import orc
    ncpu = os.cpu_count() or 1
    g, goffs, cat, roffs, _ = make_workload(args, 0)
    t0 = time.time()
    oi = orc.Index.build(g, goffs, ["chr8"], w=W, k=K, threads=ncpu)
    t_build = time.time() - t0
    sample = min(args.reads, args.ref_sample)
    names = ["r%06d" % i for i in range(sample)]
    opts = orc.AlignOpts.default(W, K)
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
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64/i32",
        "data": "synthetic", "config": workload_config(args, world),
        "cpu_baseline": {"value": val, "unit": "bases/s", "cores": ncpu, "kind": "port",
                         "sample": "%d reads x %d bp per step (of %d), %d host threads; index built by the same CPU port in %.1f s (%.4f Gbp/s)"
                                   % (sample, args.read_len, args.reads, ncpu, t_build, g.size / t_build / 1e9)},
        "e2e": {"value": val, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "index_build": {"seconds": t_build, "gbp_per_s": g.size / t_build / 1e9, "threads": ncpu},
    }
    print(json.dumps(line))


def workload_config(args, world):
    return {"workload": "BASELINE configs[1]: synthetic %.1f Mbp random genome, index k=%d w=%d b=14; %d simulated %d bp ONT-like reads "
                        "(%.1f%% error) per GPU per step" % ((args.genome_mbp or 145.138636), K, W, args.reads, args.read_len, 100 * sum(ERR)),
            "reads_per_gpu": args.reads, "read_len": args.read_len, "parallelism": "reads sharded over %d GPU(s), index replicated" % world,
            "l2": "inputs (%.2f GB of reads per step) are larger than the 126 MB L2" % (args.reads * args.read_len / 1e9)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--reads", type=int, default=100_000)
    ap.add_argument("--read-len", type=int, default=10_000)
    ap.add_argument("--genome-mbp", type=float, default=0.0)
    ap.add_argument("--ref-sample", type=int, default=20_000, help="reads per step of the CPU reference arm")
    ap.add_argument("--cpu-sample", type=int, default=3000, help="reads of the single-thread cpu_baseline leg")
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
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    g, goffs, cat, roffs, t_gen = make_workload(args, rank)
    n_bases = int(roffs[-1])
    stream = torch.cuda.current_stream()
    ctx = mm2.Context(local_rank, stream=stream.cuda_stream)

    # ---- index build (once per rank; reported, not part of the step) -----------------------------------------------------
    pin_g = mm2.PinnedBuffer(g.size)
    pg = pin_g.array(np.uint8, g.size)
    pg[:] = g
    t0 = time.perf_counter()
    gi = mm2.Index.build(ctx, pg, goffs, ["chr8"], w=W, k=K, b=14)
    t_build = time.perf_counter() - t0
    mm2.Index.build(ctx, pg, goffs, ["chr8"], w=W, k=K, b=14).close()   # warm-up: kernels loaded, arenas and the index pool grown
    t_build_warm = 1e9
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        gi2 = mm2.Index.build(ctx, pg, goffs, ["chr8"], w=W, k=K, b=14)   # from pinned host ASCII to an index resident in HBM
        t_build_warm = min(t_build_warm, time.perf_counter() - t0)
        bt = gi2.build_timings()
        gi2.close()
    n_keys = gi.stats()[0]

    # ---- reads: pinned host copy (e2e) and HBM-resident copy (value) ---------------------------------------------------------
    pin_r = mm2.PinnedBuffer(cat.size)
    pr = pin_r.array(np.uint8, cat.size)
    pr[:] = cat
    d_cat = torch.empty(cat.size + 64, dtype=torch.uint8, device="cuda")
    d_cat[:cat.size].copy_(torch.from_numpy(cat))
    d_off = torch.from_numpy(roffs.astype(np.int64)).cuda()
    opts = mm2.default_map_opts(W, K)
    torch.cuda.synchronize()

    def step_dev():
        return ctx.map_batch(gi, None, roffs, opts, device_ptrs=(d_cat.data_ptr(), d_off.data_ptr()))

    def step_e2e():
        return ctx.map_batch(gi, pr, roffs, opts)

    for _ in range(args.warmup):
        step_dev().close()
    # ---- timed region: device-resident -----------------------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.launch_count
    stage_ms = {}
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        res = step_dev()
        for kname, v in ctx.last_timings().items():
            stage_ms[kname] = stage_ms.get(kname, 0.0) + v / args.steps
        stats = dict(res.stats)
        n_recs = res.n_recs
        res.close()
    ev1.record(stream)
    barrier()
    dev_ms = ev0.elapsed_time(ev1) / args.steps
    launches = (ctx.launch_count - l0) // args.steps
    # ---- timed region: end to end from pinned host memory ------------------------------------------------------------------------
    step_e2e().close()
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(args.steps):
        res = step_e2e()
        d2h = args.reads * 64 + 8
        res.close()
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / args.steps
    barrier()
    clocks = sampler.stop()

    # ---- max over ranks ------------------------------------------------------------------------------------------------------------
    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(n_bases)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    total_bases = float(tot[0])

    if rank == 0:
        peak, peak_src = peaks()
        nm, na = stats["n_minimizers"], stats["n_anchors"]
        # algorithmic bytes per launch (SURVEY.md §8d / DESIGN.md): sketch L + 16 n_min; lookup 16 n_min + 16 n_min; anchors 8 per occurrence + sort 32 n_anchor
        # (the lookup stage now ends at the compact hit lists; the anchors are built, sorted and written once by the sort stage)
        alg = {"sketch": n_bases + 16 * nm, "lookup": 32 * nm, "anchor_sort": 8 * na + 32 * na}
        kernels = {}
        # device stages (CUDA-event timers on the launching stream) vs host wall-clock entries; shares are of the device stages
        dev_stages = {kn: ms for kn, ms in stage_ms.items() if not kn.startswith("host_") and kn not in ("h2d", "d2h", "end")}
        for kname, ms in stage_ms.items():
            e = {"ms": ms}
            if kname in dev_stages:
                e["share"] = ms / max(1e-9, sum(dev_stages.values()))
            if kname in alg and ms > 0:
                e["algorithmic_GB"] = alg[kname] / 1e9
                e["achieved_GBps"] = alg[kname] / 1e9 / (ms / 1e3)
                e["frac_of_hbm_peak"] = e["achieved_GBps"] / peak
            kernels[kname] = e
        # DRAM traffic per unit from the ncu --set full capture of the same kernels (profiles/r01_ncu_final.md, 200 Mbase
        # launch: sketch 0.744 GB / 200 Mbase, chain 1.360 GB / 13.96 M anchors, lookup 2.868 GB / 37.26 M minimizers),
        # scaled to the units of this launch
        ncu_traffic = {k_: v_[0] / v_[1] * {"base": n_bases, "anchor": na, "minimizer": nm}[v_[2]] for k_, v_ in NCU_TRAFFIC.items()}
        dom = max(dev_stages, key=dev_stages.get)
        if dom in alg:
            roof = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["achieved_GBps"], "peak": peak, "unit": "GB/s",
                    "frac": kernels[dom]["achieved_GBps"] / peak, "traffic": ncu_traffic.get(dom), "peak_source": peak_src,
                    "note": NCU_NOTE.get(dom, "") + "; traffic scaled from the ncu capture to the units of this launch"}
        else:  # chaining: integer-pipe / latency bound; credited with its HBM-visible algorithmic traffic (anchors + DP state)
            chain_bytes = na * (16 + 32 + 32 + 4)
            ach = chain_bytes / 1e9 / (stage_ms[dom] / 1e3)
            roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": ncu_traffic.get("chain"),
                    "peak_source": peak_src, "note": NCU_NOTE["chain"] + "; algorithmic bytes = 84 B per anchor; traffic scaled from the ncu capture"}
        line = {
            "metric": "mapped_bases_per_sec", "value": total_bases / (dev_ms / 1e3), "unit": "bases/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64/i32", "data": "synthetic", "config": workload_config(args, world),
            "e2e": {"value": total_bases / (e2e_ms / 1e3), "unit": "bases/s", "h2d_bytes_per_step": int(cat.size + roffs.size * 8),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms, "timing": "wall clock around mm2_map_batch (pinned host buffers)"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "kernels": kernels,
            "index_build": {"genome_bp": int(g.size), "wall_s_first": t_build, "wall_s_warm": t_build_warm,
                            "gbp_per_s_warm": g.size / t_build_warm / 1e9, "device_ms": bt, "n_keys": int(n_keys)},
            "work": {"reads": args.reads * world, "bases_per_step": total_bases, "minimizers": int(nm), "anchors": int(na), "paf_records": n_recs,
                     "rescued": int(stats["n_rescued"])},
        }
        if not args.no_cpu_baseline and world == 1:   # the CPU leg is an N = 1 measurement (rank 0's host cores are shared with the other ranks otherwise)
            from oracle import orc
            t0 = time.time()
            oi = orc.Index.build(g, goffs, ["chr8"], w=W, k=K, threads=os.cpu_count() or 1)
            t_ob = time.time() - t0
            sample = min(args.reads, args.cpu_sample)
            names = ["r%06d" % i for i in range(sample)]
            lines_cpu, st = oi.align_batch(cat[:int(roffs[sample])], roffs[:sample + 1], names, orc.AlignOpts.default(W, K), threads=1)
            res = ctx.map_batch(gi, cat[:int(roffs[sample])], roffs[:sample + 1], opts)
            same = res.paf_lines(names) == lines_cpu
            res.close()
            line["cpu_baseline"] = {"value": int(roffs[sample]) / st.seconds, "unit": "bases/s", "cores": 1, "kind": "port",
                                    "sample": "first %d reads of the step on ONE host thread (what `mm2rs align` does per read); "
                                              "CPU index build with %d threads took %.1f s" % (sample, os.cpu_count() or 1, t_ob),
                                    "paf_identical_to_gpu_on_sample": bool(same), "cells_per_anchor": st.cells / max(1, st.n_anchors)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
