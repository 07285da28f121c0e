#!/bin/bash
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest.log); tail -4 gpurun_out/r02_pytest.log | cut -c1-200
: > gpurun_out/r02_flat_ab.log
for rep in 1 2; do for lib in experiments/ab/*.so; do
  echo "== $lib" >> gpurun_out/r02_flat_ab.log
  TIDK_CUDA_LIB=$PWD/$lib timeout 300 python tools/prof_flat.py 2>&1 | grep "flat=1" | grep -E "W= 10000|W=  1000" >> gpurun_out/r02_flat_ab.log
done; done
cut -c1-110 gpurun_out/r02_flat_ab.log
TIDK_C5_CHECK_ALL=1 timeout 1500 python tests/parity_tools/c5_full.py > gpurun_out/r02_c5_full.log 2>&1; tail -5 gpurun_out/r02_c5_full.log | cut -c1-300
export TIDK_PROF_C4_SCALE=1.0 TIDK_PROF_READS=2000
timeout 400 python tools/prof_r02.py > gpurun_out/r02_prof_c4full_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'window_counts_flat_kernel|planes_build' -c 6 -o gpurun_out/r02_prof_c4full python tools/prof_r02.py > gpurun_out/r02_ncu_c4full.log 2>&1
echo "c4 full capture rc=$?"; grep "C4" gpurun_out/r02_prof_c4full_plain.log
