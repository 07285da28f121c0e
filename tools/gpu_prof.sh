#!/bin/bash
# second GPU call: failing test re-run, launch list of a short bench, full ncu capture of the hot kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "host_packed or third_party or multi_device" > gpurun_out/r02_pytest2.log 2>&1; tail -3 gpurun_out/r02_pytest2.log
export TIDK_BENCH_PASSES=4
timeout 500 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_short.json 2> gpurun_out/r02_bench_short.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/r02_ncu_launches.log 2>&1
echo "launch list rc=$?"
unset TIDK_BENCH_PASSES
timeout 300 python tools/prof_r02.py > gpurun_out/r02_prof_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'explore_vertical_kernel|window_counts_kernel|explore_planes_kernel|planes_build' -c 24 -o gpurun_out/r02_prof python tools/prof_r02.py > gpurun_out/r02_ncu_full.log 2>&1
echo "full capture rc=$?"
cat gpurun_out/r02_prof_plain.log
