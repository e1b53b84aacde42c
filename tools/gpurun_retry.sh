#!/bin/bash
# usage: tools/gpurun_retry.sh <gpus> <timeout> <command...> : retries while the pod answers "transient" (nothing charged)
gpus=$1; to=$2; shift 2
for attempt in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun --gpus $gpus --timeout $to -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|status=busy\|exit code 3"; then sleep 120; continue; fi
  echo "$out"; exit 0
done
echo "$out"; echo "gave up after 12 attempts"
