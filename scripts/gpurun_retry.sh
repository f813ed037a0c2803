#!/bin/bash
# usage: scripts/gpurun_retry.sh <tag> <timeout> [--gpus N] <command>   -- retries while the pod answers busy
tag=$1; shift; to=$1; shift
extra=""
if [ "$1" = "--gpus" ]; then extra="--gpus $2"; shift; shift; fi
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $to $extra -- "$@" > gpurun_out/$tag.gpurun.log 2>&1
  rc=$?
  if grep -q "status=transient" gpurun_out/$tag.gpurun.log; then sleep 120; continue; fi
  break
done
echo "done rc=$rc" >> gpurun_out/$tag.gpurun.log
