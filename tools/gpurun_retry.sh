#!/bin/bash
# usage: [GPUS=N] tools/gpurun_retry.sh <log> <timeout> <command...>   -- retries while the pod answers "busy" (nothing charged)
log=$1; shift; to=$1; shift
gpus=""; [ -n "$GPUS" ] && gpus="--gpus $GPUS"
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun $gpus --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient" "$log" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  exit $rc
done
exit 3
