#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
DEVS=$(seq -s, 0 $((N-1)))
TIDK_TEST_DEVICES=$DEVS timeout 600 python -m pytest tests/test_gpu_multi_device.py -m gpu -x -q 2>&1 | tail -2
timeout 900 python tools/multi_gpu.py $N 2>&1 | tail -8 | cut -c1-1000; cp gpurun_out/r02_multi_device.json gpurun_out/r02_multi_device_n$N.json
