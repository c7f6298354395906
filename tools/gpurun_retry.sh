#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> <command...>  — retries while the pod answers "busy" (exit code 3)
t=$1; shift
for k in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $GPURUN_ARGS --timeout $t -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
