#!/bin/bash
# usage: tools/gpurun_retry.sh <gpus> <timeout> <command...> : retries while the pod answers "transient" / busy (nothing charged)
gpus=$1; to=$2; shift 2
garg=""; [ "$gpus" != "1" ] && garg="--gpus $gpus"
for attempt in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun $garg --timeout $to -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|status=busy\|another call\|exit code 3\|exit code 2"; then sleep 60; continue; fi
  echo "$out"; exit 0
done
echo "$out"; echo "gave up after 20 attempts"
