#!/bin/bash
# retries a gpurun call until the pod takes it: gpu_retry.sh <log> <gpurun args...>
log=$1; shift
for k in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  if ! grep -q "status=transient\|exit code 3" "$log"; then exit 0; fi
  sleep 150
done
