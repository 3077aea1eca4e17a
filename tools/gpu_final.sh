#!/bin/bash
# One GPU-box session for the round's evidence: parity tests, bench (both arms), ncu launch list, ncu --set full of the mapping step.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_final.log 2>&1; tail -3 gpurun_out/pytest_final.log
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; echo "ref rc=$?"
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 300 python tools/prof_step.py --reads 10000 --builds 1 --steps 1 > gpurun_out/prof_step.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sketch_tile|chain_ring|seed_hits|anchor_msort|os_pass|os_hist|rs_scatter|lookup_build|tab_fill|run_fill|p_fill' -o gpurun_out/full_step_r2d -f python tools/prof_step.py --reads 10000 --builds 1 --steps 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/prof_step.log; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/full_step_r2d.ncu-rep
