#!/bin/bash
# tools/gpurun_retry.sh [gpurun args...] -- retries while the pod answers "busy" (exit code 3: nothing charged)
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
