#!/bin/bash
# usage: scripts/gpu_retry.sh <timeout-seconds> <log-name> [--gpus N] -- '<command>'   (retries while the pod answers "busy", rc 3)
T=$1; NAME=$2; shift 2
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" "$@" > "gpurun_out/$NAME.out" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
