#!/bin/bash
# gpurun with retries while the pod answers "busy / draining" (exit 3, nothing charged)
#   scripts/gpurun_retry.sh [--gpus N] --timeout S -- 'command'
for attempt in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
