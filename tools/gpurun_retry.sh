#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> [gpurun args...]   -- retries while the pod answers "transient/busy" (nothing charged)
log=$1; shift
for i in $(seq 1 30); do
  gpurun "$@" > "$log" 2>&1
  if grep -q "status=transient\|no box or slot\|answers busy" "$log"; then sleep 90; continue; fi
  break
done
tail -40 "$log"
