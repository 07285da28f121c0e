#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest.log); tail -3 gpurun_out/r02_pytest.log | cut -c1-200
timeout 900 python tests/parity_tools/full_parity.py > gpurun_out/r02_full_parity.log 2>&1; tail -1 gpurun_out/r02_full_parity.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n1.json'))
print({k:d[k] for k in ('value','passes_per_step','timed_region_s','kernel_ms_per_pass','wall_over_kernel')}, d['e2e']['value'], d['e2e_fasta']['value'])
ex=d['roofline_explore']; print(ex['k5_12']['scan_kernel_ms'], ex['k5_12']['roofline']['frac'], ex['k6_single']['scan_kernel_ms'], ex['e2e_host_pointer']['ms_per_call'], ex['cpu_baseline']['run_lists_equal_to_gpu'])"
