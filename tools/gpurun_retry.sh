#!/bin/bash
# gpurun with retries while the pod is busy (exit code 3: nothing charged).  usage: tools/gpurun_retry.sh <timeout> '<command>'
for attempt in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] pod busy, attempt $attempt; sleeping 120 s"
  sleep 120
done
exit 3
