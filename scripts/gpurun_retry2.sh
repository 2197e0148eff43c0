#!/bin/bash
# usage: gpurun_retry2.sh <gpus> <timeout> <command...>   -- like gpurun_retry.sh, on <gpus> GPUs of one box
G=$1; T=$2; shift; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|status=busy\|nothing was charged"; then sleep 150; continue; fi
  echo "$out" | tail -60; exit 0
done
echo "gave up"; echo "$out" | tail -5
