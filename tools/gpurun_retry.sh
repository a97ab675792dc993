#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3 / transient: nothing is charged).
# usage: tools/gpurun_retry.sh [gpurun options] -- 'command'
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  rc=$?
  if echo "$out" | grep -q "status=transient"; then
    sleep 90
    continue
  fi
  echo "$out"
  exit $rc
done
echo "gpurun_retry: no slot after 40 attempts"
exit 3
