#!/bin/bash
# usage: tools/gpurun_retry.sh LOGFILE [gpurun args...] -- retries while the pod answers "busy" (exit 3 / transient)
LOG=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if grep -q "status=transient\|no box or slot" "$LOG" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  exit $rc
done
exit 3
