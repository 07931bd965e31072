#!/bin/bash
# retry a gpurun call while the pod answers "busy" (exit code 3); usage: tools/gpurun_retry.sh [gpurun flags] -- 'command'
for attempt in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] attempt $attempt answered busy; sleeping 150 s"
  sleep 150
done
exit 3
