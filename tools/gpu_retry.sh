#!/bin/bash
# tools/gpu_retry.sh TIMEOUT 'command' -- gpurun with retries while the pod answers "busy" (exit code 3: nothing charged)
T=$1; shift
for i in $(seq 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
