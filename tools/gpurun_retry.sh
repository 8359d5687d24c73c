#!/bin/bash
# usage: tools/gpurun_retry.sh [gpurun flags ...] -- 'command'    retries while the pod answers busy (exit code 3)
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gpurun_retry] busy (attempt $attempt), sleeping 150 s" >&2
  sleep 150
done
exit 3
