#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <command string> ; retries while the pod answers "transient" / busy (nothing charged)
T=$1; shift
for i in $(seq 1 40); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  if echo "$OUT" | grep -q "status=transient\|retry in a few minutes\|no box or slot"; then
    sleep 120
    continue
  fi
  echo "$OUT"
  exit 0
done
echo "gave up after 40 tries"; echo "$OUT" | tail -5
