#!/bin/bash
# usage: tools/gpurun_retry.sh TIMEOUT 'command'   -- retries while the pod answers "transient" / busy (nothing charged)
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|nothing was charged" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  echo "$out" | tail -40; exit $rc
done
echo "gave up after 40 tries"; exit 3
