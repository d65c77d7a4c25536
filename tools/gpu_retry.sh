#!/bin/bash
# usage: [GPUS=N] tools/gpu_retry.sh <logfile> <timeout> <command...> : retries while gpurun answers "busy"
log=$1; shift; to=$1; shift
extra=""; [ -n "$GPUS" ] && extra="--gpus $GPUS"
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout $to $extra -- "$@" > $log 2>&1
  rc=$?
  if grep -q "status=transient" $log || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
echo "gpu_retry done rc=$rc" >> $log
