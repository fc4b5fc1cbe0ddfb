#!/bin/bash
# Retry a gpurun call while the pod answers "busy" (exit 3: nothing charged).  usage: tools/gpurun_retry.sh <timeout> <command...>
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" /tmp/gpurun_last.log; then break; fi
  sleep 90
done
tail -n 80 /tmp/gpurun_last.log
echo "gpurun_retry rc=$rc after $i tries"
