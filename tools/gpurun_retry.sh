#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <timeout> '<command>'  -- retries while the pod is busy
log=$1; to=$2; cmd=$3
for try in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout $to -- "$cmd" > $log 2>&1
  if grep -q "status=transient\|nothing was charged" $log; then sleep 90; continue; fi
  break
done
