#!/bin/bash
# gpurun with retries while the pod has no free GPU slot (nothing is charged for those attempts).
#   tools/grun.sh <timeout-seconds> '<command>'
T="$1"; shift
for attempt in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[grun] no slot (attempt $attempt), retrying in 60 s" >&2
  sleep 60
done
exit 3
