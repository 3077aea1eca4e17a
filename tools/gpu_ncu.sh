#!/bin/bash
# ncu evidence: launch list of the bench command + one --set full capture of the mapping kernels (tools/prof_step.py)
set -x
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'sketch_tile|chain_ring|seed_hits|anchor_msort|rs_scatter|tab_build' -o gpurun_out/full_step -f python tools/prof_step.py --builds 1 --steps 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
