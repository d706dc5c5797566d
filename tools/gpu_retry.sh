#!/bin/bash
# tools/gpu_retry.sh TIMEOUT 'command' -- gpurun with retries while the pod answers "busy" (exit code 3: nothing charged)
T=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$T" "${@:1:$#-1}" -- "${@: -1}"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
