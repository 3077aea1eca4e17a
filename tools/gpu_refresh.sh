#!/bin/bash
# One GPU-box session: parity tests, bench (both arms), ncu launch list, ncu --set full of the mapping step.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 300 python tools/prof_step.py --builds 1 --steps 2 > gpurun_out/prof_step.log 2>&1; echo "prof rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sketch_tile|chain_ring|chain_classify|lookup_count|anchor_msort|anchor_fill|filter_kernel|rs_scatter|rs_hist' -o gpurun_out/full_step -f python tools/prof_step.py --builds 1 --steps 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json
