#!/bin/bash
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest.log); tail -4 gpurun_out/r02_pytest.log | cut -c1-200
: > gpurun_out/r02_span_ab.log
for rep in 1 2; do for v in 8 4 2; do
  echo "== span$v" >> gpurun_out/r02_span_ab.log
  TIDK_CUDA_LIB=$PWD/experiments/ab/libtidk_span$v.so timeout 300 python tools/prof_span.py 0.125 0.5 >> gpurun_out/r02_span_ab.log 2>&1
done; done
TIDK_CUDA_LIB=$PWD/experiments/ab/libtidk_span8.so timeout 300 python tools/prof_span.py 1.0 >> gpurun_out/r02_span_ab.log 2>&1
TIDK_CUDA_LIB=$PWD/experiments/ab/libtidk_span4.so timeout 300 python tools/prof_span.py 1.0 >> gpurun_out/r02_span_ab.log 2>&1
cat gpurun_out/r02_span_ab.log | cut -c1-220
timeout 300 python tools/multi_gpu.py 1 2>&1 | tail -3 | cut -c1-600
