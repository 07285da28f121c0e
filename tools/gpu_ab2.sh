#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_window_counts.py tests/test_third_party_vectors.py tests/test_gpu_multi_device.py tests/test_gpu_cli.py -m gpu -x -q > gpurun_out/r02_pytest4.log 2>&1; tail -15 gpurun_out/r02_pytest4.log | cut -c1-300
timeout 300 python tools/prof_flat.py > gpurun_out/r02_flat.log 2>&1; cat gpurun_out/r02_flat.log | cut -c1-200
TIDK_CUDA_LIB=$PWD/experiments/ab/libtidk_flat4.so timeout 300 python tools/prof_flat.py 2>&1 | grep "flat=1" > gpurun_out/r02_flat4.log; cat gpurun_out/r02_flat4.log | cut -c1-200
: > gpurun_out/r02_ab2.log
for rep in 1 2; do for lib in base sync; do
  echo "== $lib" >> gpurun_out/r02_ab2.log
  TIDK_CUDA_LIB=$PWD/experiments/ab/libtidk_$lib.so TIDK_PROF_CFG=c5 TIDK_PROF_STEPS=6 timeout 200 python tools/prof_explore.py 2>&1 | tail -3 >> gpurun_out/r02_ab2.log
done; done
cat gpurun_out/r02_ab2.log | cut -c1-120
