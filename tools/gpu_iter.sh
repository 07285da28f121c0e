#!/bin/bash
# iteration call: explore + window parity tests, timings of the hot kernels, full-size parity
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest3.log 2>&1; tail -3 gpurun_out/r02_pytest3.log
timeout 300 python tools/prof_r02.py > gpurun_out/r02_prof_plain2.log 2>&1; cat gpurun_out/r02_prof_plain2.log
timeout 900 python tests/parity_tools/full_parity.py > gpurun_out/r02_full_parity.log 2>&1; tail -2 gpurun_out/r02_full_parity.log | cut -c1-300
