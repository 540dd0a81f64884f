#!/bin/bash
# gpurun with retries on "no slot right now" (exit code 3); usage: tools/gpurun_retry.sh <timeout> [--gpus N] -- '<command>'
t=$1; shift
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$t" "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
