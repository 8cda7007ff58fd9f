#!/bin/bash
# like gpu_retry.sh but gives up after 3 attempts
log=$1; shift
for k in 1 2 3; do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  if ! grep -q "status=transient\|exit code 3" "$log"; then exit 0; fi
  sleep 60
done
