#!/bin/bash
# usage: gpurun_retry.sh <timeout> <command...>   -- retries while the pod answers "transient"/busy (nothing charged)
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|status=busy\|nothing was charged"; then sleep 150; continue; fi
  echo "$out" | tail -60; exit 0
done
echo "gave up"; echo "$out" | tail -5
