#!/bin/bash
# usage: tools/gpu.sh <tag> <timeout_s> [--gpus N] -- '<command>'   (retries while the pod answers "busy")
tag=$1; shift; to=$1; shift
extra=()
while [ "$1" != "--" ]; do extra+=("$1"); shift; done
shift
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$to" "${extra[@]}" -- "$1" > "gpurun_out/call_$tag.log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "gpurun_out/call_$tag.log"; then break; fi
  sleep 45
done
echo "rc=$rc" >> "gpurun_out/call_$tag.log"
