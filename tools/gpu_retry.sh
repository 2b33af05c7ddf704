#!/bin/bash
# gpu_retry.sh TIMEOUT CMD...: gpurun, retried every two minutes while the pod answers "busy" (exit code 3)
t=$1; shift
for i in $(seq 1 15); do
  /usr/local/graft/bin/gpurun --timeout $t -- "$@" > /tmp/gpu_retry.out 2>&1; rc=$?
  if grep -q "status=transient" /tmp/gpu_retry.out || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
tail -40 /tmp/gpu_retry.out
