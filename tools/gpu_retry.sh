#!/bin/bash
# usage: tools/gpu_retry.sh [--gpus N] TIMEOUT_SECONDS 'command'   -- retries while the pod answers busy / transient
GP=""
if [ "$1" == "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun $GP --timeout "$T" -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if [ $rc -eq 3 ] || grep -q "status=transient" /tmp/gpurun_last.log; then
    sleep 120; continue
  fi
  break
done
cat /tmp/gpurun_last.log
exit $rc
