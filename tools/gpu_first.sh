#!/bin/bash
# first GPU call of a session: tests, bench line, reference arm (run on the GPU box via gpurun)
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest.log)
timeout 500 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
tail -3 gpurun_out/r02_pytest.log; tail -c 1500 gpurun_out/r02_bench_n1.err
