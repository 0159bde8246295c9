#!/bin/bash
# scratch: gpurun with retries while the pod answers busy (exit code 3: nothing charged).  usage: gpurun_retry.sh <timeout> '<command>'
for k in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 45
done
exit 3
