#!/bin/bash
# A/B of kernel-tuning builds (experiments/ab/*.so, loaded through TIDK_CUDA_LIB), same box, alternating:
#   TIDK_AB_MODE=flat     flat window counter on shares of the C4 genome (tools/prof_span.py)
#   TIDK_AB_MODE=explore  C5-shaped explore scan (tools/prof_explore.py)
# TIDK_AB_TEST="lib.so ...": parity tests with these builds afterwards
mkdir -p gpurun_out; : > gpurun_out/r02_ab.log
for rep in 1 2; do for lib in experiments/ab/*.so; do
  echo "== $lib" >> gpurun_out/r02_ab.log
  if [ "${TIDK_AB_MODE:-flat}" = "flat" ]; then
    TIDK_CUDA_LIB=$PWD/$lib timeout 300 python tools/prof_span.py 0.125 1.0 >> gpurun_out/r02_ab.log 2>&1
  else
    TIDK_CUDA_LIB=$PWD/$lib TIDK_PROF_CFG=c5 TIDK_PROF_STEPS=6 timeout 200 python tools/prof_explore.py 2>&1 | tail -3 >> gpurun_out/r02_ab.log
  fi
done; done
cut -c1-200 gpurun_out/r02_ab.log
for lib in $TIDK_AB_TEST; do TIDK_CUDA_LIB=$PWD/experiments/ab/$lib timeout 600 python -m pytest tests/test_gpu_window_counts.py tests/test_gpu_explore.py tests/test_gpu_multi_device.py -m gpu -x -q 2>&1 | tail -2; done
