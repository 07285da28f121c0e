#!/bin/bash
# A/B of kernel-tuning builds (experiments/ab/*.so, loaded through TIDK_CUDA_LIB): C5-shaped explore scan, same box, alternating
mkdir -p gpurun_out; : > gpurun_out/r02_ab.log
for rep in 1 2; do for lib in experiments/ab/*.so; do
  echo "== $lib" >> gpurun_out/r02_ab.log
  TIDK_CUDA_LIB=$PWD/$lib TIDK_PROF_CFG=c5 TIDK_PROF_STEPS=6 timeout 200 python tools/prof_explore.py 2>&1 | tail -3 >> gpurun_out/r02_ab.log
  TIDK_CUDA_LIB=$PWD/$lib TIDK_PROF_CFG=c5 TIDK_PROF_STEPS=4 TIDK_PROF_KMIN=6 TIDK_PROF_KMAX=6 timeout 200 python tools/prof_explore.py 2>&1 | tail -1 >> gpurun_out/r02_ab.log
done; done
cut -c1-150 gpurun_out/r02_ab.log
for lib in $TIDK_AB_TEST; do TIDK_CUDA_LIB=$PWD/experiments/ab/$lib timeout 600 python -m pytest tests/test_gpu_explore.py -m gpu -x -q 2>&1 | tail -2; done
