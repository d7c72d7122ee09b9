#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout-seconds> [--gpus N] '<command>'   (build container only: retries while the pod is busy)
T=$1; shift
G=""
if [ "$1" = "--gpus" ]; then G="--gpus $2"; shift 2; fi
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" $G -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
