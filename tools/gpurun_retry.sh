#!/bin/bash
# gpurun with retries while the pod answers "busy / draining" (exit code 3: nothing charged)
#   bash tools/gpurun_retry.sh <timeout_s> '<command>' [log]
T=$1; CMD=$2; LOG=${3:-/tmp/gpurun_retry.log}
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$CMD" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $LOG; then break; fi
  sleep 90
done
tail -40 $LOG
exit $rc
