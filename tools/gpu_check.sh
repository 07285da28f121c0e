#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r02_bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n1.json'))
print({k:d[k] for k in ('value','steps','passes_per_step','timed_region_s','ms_per_step','kernel_ms_per_pass','wall_over_kernel','gpu_launches')}, d['e2e']['value'], d['e2e_fasta']['value'], d['e2e_fasta'].get('devices_used'))"
timeout 1500 python tests/parity_tools/stress.py 20000 3000 > gpurun_out/r02_stress.log 2>&1; echo "stress rc=$?"; tail -3 gpurun_out/r02_stress.log
