#!/usr/bin/env bash
# retry a gpurun call while the pod answers "busy" (exit 3 / transient); usage: gpurun_retry.sh <timeout> <cmd>
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "$2" > /tmp/gpurun_last.log 2>&1
  if ! grep -q "status=transient" /tmp/gpurun_last.log; then break; fi
  sleep 100
done
tail -5 /tmp/gpurun_last.log
