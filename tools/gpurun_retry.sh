#!/bin/bash
# usage: [GPUS=n] gpurun_retry.sh <timeout-seconds> <command...>   — retries while the pod answers "busy" (exit code 3)
t=$1; shift
extra=""
if [ -n "$GPUS" ]; then extra="--gpus $GPUS"; fi
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $extra --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
