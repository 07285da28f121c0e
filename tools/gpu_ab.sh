#!/bin/bash
# A/B of kernel-tuning builds (experiments/ab/*.so, loaded through TIDK_CUDA_LIB): flat window counter, shares of the C4 genome
mkdir -p gpurun_out; : > gpurun_out/r02_ab.log
for rep in 1 2; do for lib in experiments/ab/*.so; do
  echo "== $lib" >> gpurun_out/r02_ab.log
  TIDK_CUDA_LIB=$PWD/$lib timeout 300 python tools/prof_span.py 0.125 1.0 >> gpurun_out/r02_ab.log 2>&1
done; done
cut -c1-200 gpurun_out/r02_ab.log
for lib in $TIDK_AB_TEST; do TIDK_CUDA_LIB=$PWD/experiments/ab/$lib timeout 600 python -m pytest tests/test_gpu_window_counts.py tests/test_gpu_multi_device.py -m gpu -x -q 2>&1 | tail -2; done
