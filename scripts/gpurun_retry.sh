#!/bin/bash
# usage: scripts/gpurun_retry.sh LOGFILE TIMEOUT [--gpus N] -- 'command'   (retries while gpurun answers 3 = no slot free)
log=$1; shift; to=$1; shift
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $to "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
