#!/bin/bash
# gpurun with retries while the pod answers "transient" (nothing charged): tools/gpurun_retry.sh <gpurun args...>
for i in 1 2 3 4 5 6 7 8 9 10; do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 120; continue; fi
  echo "$out"; exit 0
done
echo "$out"
