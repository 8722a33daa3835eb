#!/bin/bash
# gpurun with retries while the pod answers busy (exit 3): tools/gpurun_retry.sh [gpurun args...] -- 'command'
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
