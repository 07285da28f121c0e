#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_cli.py -m gpu -x -q 2>&1 | tail -2
timeout 1500 python tests/parity_tools/stress.py 5000 250 > gpurun_out/r02_stress.log 2>&1; echo "stress rc=$?"; tail -5 gpurun_out/r02_stress.log
