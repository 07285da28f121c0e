#!/bin/bash
# usage: tools/gpurun_retry_n.sh NGPUS TIMEOUT 'command'
N=$1; T=$2; shift 2
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --gpus "$N" --timeout "$T" "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|nothing was charged" || [ $rc -eq 3 ]; then sleep 120; continue; fi
  echo "$out" | tail -60; exit $rc
done
echo "gave up"; exit 3
