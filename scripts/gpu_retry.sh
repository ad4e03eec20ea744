#!/bin/bash
# usage: scripts/gpu_retry.sh <out-file> <timeout-seconds> <command...>   — retries while the pod answers "transient" / busy (exit 3)
out=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout $to -- "$@" > $out 2>&1
  rc=$?
  if grep -q "status=transient" $out || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
echo "done rc=$rc" >> $out
