#!/bin/bash
# multi-GPU call: usage  gpurun --gpus N -- 'bash tools/gpu_multi.sh N'
N=${1:-2}
mkdir -p gpurun_out
DEVS=$(seq -s, 0 $((N-1)))
TIDK_TEST_DEVICES=$DEVS timeout 600 python -m pytest tests/test_gpu_multi_device.py -m gpu -x -q 2>&1 | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench N=$N rc=$?"; tail -c 600 gpurun_out/r02_bench_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n$N.json'))
print({k:d[k] for k in ('value','n_gpus','ms_per_step','kernel_ms_per_pass','wall_over_kernel','kernel_only_value')}, d['e2e']['value'], d.get('e2e_fasta',{}).get('value'))"
timeout 900 python tools/multi_gpu.py $N 2>&1 | tail -8 | cut -c1-900; cp gpurun_out/r02_multi_device.json gpurun_out/r02_multi_device_n$N.json
