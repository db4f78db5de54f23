#!/bin/bash
# gpurun with retries on "no box / slot free" (exit code 3, nothing charged).  usage: tools/grun.sh [gpurun options] -- 'command'
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[grun] busy (attempt $i), retrying in 120 s" >&2
  sleep 120
done
exit 3
