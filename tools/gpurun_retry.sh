#!/bin/bash
# usage: tools/gpurun_retry.sh [--gpus N] <timeout> <command...>   -- retries while the pod answers "transient" (busy), max ~40 min
GP=""
if [ "$1" == "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun $GP --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 100; continue; fi
  echo "$out"; exit 0
done
echo "$out"; echo "gave up after retries"
