#!/bin/bash
# usage: tools/gpurun_retry.sh [--gpus N] [--timeout S] -- 'command'   (retries while the pod is busy; rc 3 = no slot)
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 20
done
exit 3
