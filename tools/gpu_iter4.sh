#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest.log); tail -4 gpurun_out/r02_pytest.log | cut -c1-200
: > gpurun_out/r02_zero_ab.log
for rep in 1 2; do
  echo "== base" >> gpurun_out/r02_zero_ab.log
  TIDK_CUDA_LIB=$PWD/experiments/ab/libtidk_base.so timeout 300 python tools/prof_span.py 0.125 0.5 >> gpurun_out/r02_zero_ab.log 2>&1
  echo "== zero-in-kernel" >> gpurun_out/r02_zero_ab.log
  timeout 300 python tools/prof_span.py 0.125 0.5 >> gpurun_out/r02_zero_ab.log 2>&1
done
cat gpurun_out/r02_zero_ab.log | cut -c1-220
