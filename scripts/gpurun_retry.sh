#!/bin/bash
# usage: [GPUS=N] scripts/gpurun_retry.sh <log> <timeout> <cmd...>   -- retries while gpurun answers "busy" (exit 3)
log=$1; shift; to=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --gpus ${GPUS:-1} --timeout $to -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "rc=$rc after $i tries" >> "$log"; exit $rc; fi
  sleep 150
done
