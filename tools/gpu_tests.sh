#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest.log); tail -6 gpurun_out/r02_pytest.log | cut -c1-200
timeout 300 python tools/multi_gpu.py 1 2>&1 | tail -2 | cut -c1-900
