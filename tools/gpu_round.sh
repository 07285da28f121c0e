#!/bin/bash
# full GPU pass of the current build: tests, bench line, reference arm, launch list, full ncu capture of the hot kernels
mkdir -p gpurun_out; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest.log); tail -4 gpurun_out/r02_pytest.log | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
export TIDK_BENCH_PASSES=4
timeout 500 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_short.json 2> gpurun_out/r02_bench_short.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/r02_ncu_launches.log 2>&1
echo "launch list rc=$?"
unset TIDK_BENCH_PASSES
timeout 300 python tools/prof_r02.py > gpurun_out/r02_prof_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none -k regex:"explore_vertical_kernel|window_counts_kernel|window_counts_flat_kernel|explore_planes_kernel|planes_build" -c 26 -o gpurun_out/r02_prof python tools/prof_r02.py > gpurun_out/r02_ncu_full.log 2>&1
echo "full capture rc=$?"
cat gpurun_out/r02_prof_plain.log
