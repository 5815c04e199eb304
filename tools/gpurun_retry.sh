#!/bin/bash
# usage: tools/gpurun_retry.sh TIMEOUT 'command' [gpurun args...]   -- retries while the pod answers "transient" (nothing charged)
t=$1; cmd=$2; shift 2
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun "$@" --timeout $t -- "$cmd" 2>&1)
  echo "$out" | tail -80
  if echo "$out" | grep -q "status=transient"; then echo "[retry $i] transient, sleeping 120 s"; sleep 120; else break; fi
done
