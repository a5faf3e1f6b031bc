#!/bin/bash
# gpurun with retries on "busy" (exit 3): bash tools/gpurun_retry.sh TIMEOUT_S [--gpus N] -- 'command'
T=$1; shift
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout $T "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] attempt $attempt answered busy; sleeping 150 s"
  sleep 150
done
exit 3
