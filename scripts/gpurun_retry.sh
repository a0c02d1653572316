#!/bin/bash
# gpurun with retries while the pod answers "busy" (status=transient, nothing charged).  usage: gpurun_retry.sh <timeout> <command string>
for attempt in 1 2 3 4 5 6 7 8 9 10; do
  out=$(/usr/local/graft/bin/gpurun --timeout "$1" -- "$2" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
