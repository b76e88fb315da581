#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <command...>   -- retries transient "no slot" answers (nothing is charged for those)
T=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10; do
  OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  echo "$OUT" | tail -${TAIL:-60}
  if echo "$OUT" | grep -q "status=transient"; then sleep 90; continue; fi
  break
done
