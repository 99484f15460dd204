#!/bin/bash
# usage: tools/gpurun_retry.sh LOG [gpurun args...] -- retries while the pod answers busy (exit 3 / transient), up to 12 times
log=$1; shift
for attempt in $(seq 1 60); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if ! grep -q "status=transient\|already running" "$log" && [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
