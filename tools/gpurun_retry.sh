#!/bin/bash
# usage: tools/gpurun_retry.sh LOGFILE [gpurun args...] -- 'command'
# retries while the pod answers busy (exit 3), up to 12 times 2 minutes apart
log=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "rc=$rc" >> "$log"; exit $rc; fi
  sleep 120
done
echo "rc=3 (gave up)" >> "$log"
exit 3
