#!/bin/bash
# torchrun bench on N GPUs of one box (strong scaling of the C4 genome) + the reference arm launched the same way
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench N=$N rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n$N.json'))
print({k:d[k] for k in ('value','n_gpus','ms_per_step','kernel_ms_per_pass','wall_over_kernel','kernel_only_value')}, d['e2e']['value'], d.get('e2e_fasta',{}).get('value'))"
