#!/bin/bash
# usage: scripts/gpurun_retry.sh <tag> <timeout> <command...>   -- retries while the pod answers busy (exit code 3)
tag=$1; shift; to=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > gpurun_out/$tag.gpurun.log 2>&1
  rc=$?
  if grep -q "status=transient" gpurun_out/$tag.gpurun.log; then sleep 120; continue; fi
  break
done
echo "done rc=$rc" >> gpurun_out/$tag.gpurun.log
