#!/bin/bash
mkdir -p gpurun_out
TIDK_CUDA_LIB=$PWD/experiments/ab/libtidk_tma.so timeout 300 python -m pytest tests/test_gpu_explore.py -m gpu -x -q 2>&1 | tail -3
TIDK_AB_MODE=explore timeout 600 bash tools/gpu_ab.sh 2>&1 | tail -14
TIDK_CUDA_LIB=$PWD/experiments/ab/libtidk_tma.so timeout 300 python -m pytest tests/test_gpu_multi_device.py tests/test_gpu_cli.py -m gpu -x -q 2>&1 | tail -2
TIDK_CUDA_LIB=$PWD/experiments/ab/libtidk_tma.so timeout 300 python tests/parity_tools/stress.py 30000 400 2>&1 | tail -2
